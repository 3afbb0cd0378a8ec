#!/bin/bash
# whole GPU suite, then the default bench line
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_c4_v4.json 2> gpurun_out/bench_c4_v4.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_v4.json')); print(d['ms_per_step'], d['value']/1e9, d['roofline']['kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['whole_path']['frac'], d['e2e'], d['clocks'], d['gpu_launches'], d['memory_gb'])"
tail -3 gpurun_out/bench_c4_v4.err
