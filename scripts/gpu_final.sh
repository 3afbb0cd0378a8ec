#!/bin/bash
# Final single-GPU evidence: default bench line, reference arm, DRAM bytes of the bucketed kernels (index form).
set -u
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_c4_final.json 2> gpurun_out/bench_c4_final.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_final.json')); print(d['ms_per_step'], d['value']/1e9, d['roofline']['kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['whole_path']['frac'], d['e2e'], d['clocks'], d['gpu_launches'])"
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference exit $?"; cat gpurun_out/bench_reference.json
CMD="python bench.py --workload c4 --steps 2 --warmup 1 --sample-reads 2000"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'sk_bucket_kernel|sk_scatter_reads|sk_scatter_buckets_kernel' -s 3 -c 3 --csv \
    --log-file gpurun_out/dram_c4.csv $CMD > gpurun_out/ncu_dram_c4.log 2>&1
echo "dram exit $?"
timeout 600 python bench.py --workload c2 > gpurun_out/bench_c2_final.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_final.json')); print('c2', d['ms_per_step'], d['value']/1e9, d['e2e'])"
timeout 600 python bench.py --workload c3 > gpurun_out/bench_c3_final.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_c3_final.json')); print('c3', d['ms_per_step'], d['value']/1e9, d['e2e'])"
