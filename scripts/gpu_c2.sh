#!/bin/bash
for t in 16384 8192 4096 2048; do
GA_SK_TARGET=$t timeout 300 python bench.py --workload c2 --sample-reads 2000 > gpurun_out/bench_c2_t.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_t.json')); print($t, round(d['ms_per_step'],2), {a:round(b,2) for a,b in d['roofline']['kernel_ms_per_step'].items()}, round(d['e2e']['ms_per_step'],1))"
done
