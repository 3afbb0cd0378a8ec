#!/bin/bash
# Round-2 bench lines of the small configurations (C1, C2, C3) and of the -c (CountMinSketch) route.
set -u
mkdir -p gpurun_out
for spec in "c1" "c2" "c3" "c2 --sketch" "c1 --sketch" "c3 --sketch"; do
  name=$(echo $spec | tr -d ' -')
  timeout 900 python bench.py --workload $spec --steps 5 --warmup 3 > gpurun_out/bench_r2_$name.json 2> gpurun_out/bench_r2_$name.err
  echo "$spec exit $?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_r2_$name.json')); k=d['roofline']['kernel_ms_per_step']
print('  step %.2f ms  e2e %.2f ms  whole-path frac %.3f  dominant %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['whole_path']['frac'], d['roofline']['kernel']))
print('  kernels', {a:round(b,2) for a,b in k.items()})
print('  user_api', d['e2e'].get('user_api'))
print('  cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])" || tail -5 gpurun_out/bench_r2_$name.err
done
