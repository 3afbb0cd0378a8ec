#!/bin/bash
# Last single-GPU visit of the round: smoke(), the bucket / exchange tests on the final build, and ncu rows for the -c
# (CountMinSketch) route on C2 and for the paired route on C3 (launch list + DRAM bytes of their kernels).
set -u
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/smoke.log
timeout 900 python -m pytest tests -q -m gpu -x -k "sources or push_kernels or segments or determin or c4_shaped or sketch or readme" > gpurun_out/last_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/last_tests.log
for spec in "c2 --sketch" "c3"; do
  name=$(echo $spec | tr -d ' -')
  CMD="python bench.py --workload $spec --steps 2 --warmup 1 --sample-reads 2000"
  GA_BENCH_SKIP_E2E=1 $CMD > gpurun_out/plain_$name.json 2> gpurun_out/plain_$name.err; echo "plain $name exit $?"
  GA_BENCH_SKIP_E2E=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -c 400 --csv --log-file gpurun_out/ncu_$name.csv $CMD > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name exit $?"
done
