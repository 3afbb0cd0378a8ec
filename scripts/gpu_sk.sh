#!/bin/bash
# GPU visit for the bucketed (super-k-mer) path: its parity tests first, then the whole GPU suite, then benches.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider -k "bucketed" > gpurun_out/pytest_sk.log 2>&1
echo "pytest bucketed exit $?" | tee -a gpurun_out/pytest_sk.log
tail -25 gpurun_out/pytest_sk.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short --durations=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest all exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for wl in c2 c4; do
  GA_TRACE=0 timeout 900 python bench.py --workload $wl --steps 3 --warmup 2 > gpurun_out/bench_${wl}_sk.json 2> gpurun_out/bench_${wl}_sk.err
  echo "bench $wl exit $?"; cat gpurun_out/bench_${wl}_sk.json; tail -5 gpurun_out/bench_${wl}_sk.err
done
