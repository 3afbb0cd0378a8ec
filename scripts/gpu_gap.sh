#!/bin/bash
set -u
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 600 python bench.py --workload c4 --steps 3 --warmup 2 --sample-reads 2000 > gpurun_out/bench_gap.json 2>gpurun_out/bench_gap.err; python -c "
import json; d=json.load(open('gpurun_out/bench_gap.json')); k=d['roofline']['kernel_ms_per_step']; print(d['ms_per_step'], sum(k.values()), d['e2e']['ms_per_step'], d['memory_gb']); print({a:round(b,2) for a,b in d['roofline']['stage_ms_per_step'].items()})"
done
