#!/bin/bash
# Round evidence on one B200: plain bench (exit 0) -> ncu launch list of the same command -> DRAM bytes of the
# bucketed kernels on the full workload -> full-set capture on a quarter-size instance with the same coverage.
set -u
mkdir -p gpurun_out
CMD="python bench.py --workload c4 --steps 2 --warmup 1 --sample-reads 2000"
$CMD > gpurun_out/plain_c4.json 2> gpurun_out/plain_c4.err
echo "plain exit $?"; python -c "
import json; d=json.load(open('gpurun_out/plain_c4.json')); print(d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['e2e'])"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_c4.csv $CMD > gpurun_out/ncu_launch_c4.log 2>&1
echo "launch list exit $?"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'sk_bucket_kernel|sk_scatter_reads|sk_scatter_buckets_kernel' -s 3 -c 3 --csv \
    --log-file gpurun_out/dram_c4.csv $CMD > gpurun_out/ncu_dram_c4.log 2>&1
echo "dram exit $?"; tail -4 gpurun_out/dram_c4.csv | cut -c1-60,200-400
R=25000000
Q="python bench.py --workload c4 --reads $R --genome $((R / 2)) --steps 1 --warmup 1 --sample-reads 2000"
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'sk_bucket_kernel|sk_scatter_reads|sk_scatter_buckets_kernel' -s 3 -c 3 \
    -o gpurun_out/prof_sk -f $Q > gpurun_out/ncu_full_sk.log 2>&1
echo "full capture exit $?"
ncu -i gpurun_out/prof_sk.ncu-rep --page raw --csv > gpurun_out/prof_sk_raw.csv 2>/dev/null
ls -la gpurun_out/prof_sk.ncu-rep gpurun_out/prof_sk_raw.csv
