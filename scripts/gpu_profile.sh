#!/bin/bash
# ncu evidence: per-launch durations of one bench run, then a full capture of the hot kernels.
set -u
mkdir -p gpurun_out
WL=${1:-c2}
CMD="python bench.py --steps 2 --warmup 1 --workload $WL --sample-reads 2000"
$CMD > gpurun_out/plain_$WL.json 2> gpurun_out/plain_$WL.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_$WL.csv $CMD > gpurun_out/ncu_launch_$WL.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:'prefilter_update_kernel|count_candidates_kernel|build_unpaired_dna_kernel|build_paired_kernel' -s 3 -c 3 \
    -o gpurun_out/prof_$WL -f $CMD > gpurun_out/ncu_full_$WL.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out
