#!/bin/bash
# First GPU visit of a change: parity tests, a short bench, and the per-launch ncu list.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short --durations=15 -p no:cacheprovider \
    > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench exit $?"; cat gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
