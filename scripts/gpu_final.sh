#!/bin/bash
# Final single-GPU evidence: default bench line, reference arm, the smaller configurations (ncu passes: gpu_prof_final.sh).
set -u
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_c4_final.json 2> gpurun_out/bench_c4_final.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_final.json')); print(d['ms_per_step'], d['value']/1e9, d['roofline']['kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['whole_path']['frac'], d['e2e'], d['clocks'], d['gpu_launches'])"
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference exit $?"; cat gpurun_out/bench_reference.json
timeout 600 python bench.py --workload c2 > gpurun_out/bench_c2_final.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_final.json')); print('c2', d['ms_per_step'], d['value']/1e9, d['e2e'])"
timeout 600 python bench.py --workload c3 > gpurun_out/bench_c3_final.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_c3_final.json')); print('c3', d['ms_per_step'], d['value']/1e9, d['e2e'])"
timeout 600 python bench.py --workload c1 > gpurun_out/bench_c1_final.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_c1_final.json')); print('c1', d['ms_per_step'], d['value']/1e9, d['e2e'])"
