#!/bin/bash
# full-set capture of the lane-per-read scatter kernel on a quarter-size C4 instance (same coverage)
set -u
mkdir -p gpurun_out
R=25000000
Q="python bench.py --workload c4 --reads $R --genome $((R / 2)) --steps 1 --warmup 1 --sample-reads 2000"
GA_BENCH_SKIP_E2E=1 $Q > gpurun_out/plain_lane_q.json 2> gpurun_out/plain_lane_q.err; echo "plain exit $?"
GA_BENCH_SKIP_E2E=1 timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'sk_scatter_reads' -s 1 -c 1 -o gpurun_out/prof_lane -f $Q > gpurun_out/ncu_full_lane.log 2>&1
echo "full capture exit $?"
ncu -i gpurun_out/prof_lane.ncu-rep --page raw --csv > gpurun_out/prof_lane_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_lane.ncu-rep --page source --csv > gpurun_out/src_lane.csv 2>/dev/null
ls -la gpurun_out/prof_lane* gpurun_out/src_lane.csv
