#!/bin/bash
# ncu evidence for the bucketed path on a reduced C4 (reads and genome scaled together: same coverage,
# same per-bucket shape): launch list + full capture.
set -u
mkdir -p gpurun_out
R=${1:-10000000}
G=$((R / 2))
CMD="python bench.py --workload c4 --reads $R --genome $G --steps 1 --warmup 1 --sample-reads 2000"
$CMD > gpurun_out/plain_sk.json 2> gpurun_out/plain_sk.err
echo "plain exit $?"; cat gpurun_out/plain_sk.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_sk.csv $CMD > gpurun_out/ncu_launch_sk.log 2>&1
echo "launch list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'sk_bucket_kernel|sk_scatter_reads_kernel|sk_scatter_buckets_kernel' -s 3 -c 3 \
    -o gpurun_out/prof_sk -f $CMD > gpurun_out/ncu_full_sk.log 2>&1
echo "full capture exit $?"
GA_TRACE=1 python bench.py --workload c4 --steps 2 --warmup 1 --sample-reads 2000 > gpurun_out/trace_c4.log 2>&1
grep trace gpurun_out/trace_c4.log | tail -40
