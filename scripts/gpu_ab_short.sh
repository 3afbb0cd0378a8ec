#!/bin/bash
# like gpu_ab.sh, one repetition, device steps only (no e2e / CPU baseline)
set -u
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  ( if [ "$envs" != "$spec" ]; then IFS=,; for e in $envs; do export "$e"; done; fi
    GA_BENCH_SKIP_E2E=1 timeout 600 python bench.py --workload c4 --steps 3 --warmup 2 --sample-reads 500 > gpurun_out/bench_ab_$name.json 2>gpurun_out/bench_ab_$name.err )
  python -c "
import json; d=json.load(open('gpurun_out/bench_ab_$name.json')); k=d['roofline']['kernel_ms_per_step']; print('$name', round(d['ms_per_step'],1), {a:round(b,1) for a,b in k.items() if b>1})"
done
