#!/bin/bash
# A/B on the same box: variants given as "name:ENV=val,ENV=val" (GA_LIB for another build of the library)
set -u
mkdir -p gpurun_out
for rep in 1 2; do
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  ( if [ "$envs" != "$spec" ]; then IFS=,; for e in $envs; do export "$e"; done; fi
    timeout 600 python bench.py --workload c4 --steps 3 --warmup 2 --sample-reads 2000 > gpurun_out/bench_ab_$name.json 2>gpurun_out/bench_ab_$name.err )
  python -c "
import json; d=json.load(open('gpurun_out/bench_ab_$name.json')); k=d['roofline']['kernel_ms_per_step']; print('$name', round(d['ms_per_step'],1), {a:round(b,1) for a,b in k.items() if b>1}, round(d['e2e']['ms_per_step'],1))"

done
done
