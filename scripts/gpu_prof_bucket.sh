#!/bin/bash
# ncu full-set capture (with source) of the bucket kernel on the quarter-size C4 instance of scripts/bucket_probe.py
set -u
mkdir -p gpurun_out
python scripts/bucket_probe.py > gpurun_out/probe_plain.log 2>&1; echo "plain exit $?"; tail -3 gpurun_out/probe_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sk_bucket_kernel -s 1 -c 1 \
    -o gpurun_out/prof_bucket_${1:-r2} -f python scripts/bucket_probe.py > gpurun_out/ncu_bucket.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_bucket.log
