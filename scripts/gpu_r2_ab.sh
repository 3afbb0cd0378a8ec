#!/bin/bash
# A/B on one B200: the bucket-path parity tests on the current build, then the C4 bench (device-resident leg only) for
# each "name:ENV=..,ENV=.." item of $AB (default: current build, previous library, lane8 scatter).
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_c4_parity.py tests/test_gpu_parity.py -q -m gpu -x \
    -k "${GA_TEST_FILTER:-c4 or bucket or segments or determin or scatter or sweep}" > gpurun_out/ab_tests.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/ab_tests.log
for item in ${AB:-new: base:GA_LIB=$PWD/genome-assembler_b200/build/libga_b200_base.so lane8:GA_SK_SCATTER=lane8 new2:}; do
  name=${item%%:*}; envs=${item#*:}
  ( for e in ${envs//,/ }; do export "$e"; done
    GA_BENCH_SKIP_E2E=1 timeout 600 python bench.py --workload c4 --steps 4 --warmup 2 --sample-reads 2000 \
      > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err )
  python -c "
import json; d=json.load(open('gpurun_out/ab_$name.json')); k=d['roofline']['kernel_ms_per_step']; print('$name', round(d['ms_per_step'],1), {a:round(b,1) for a,b in k.items() if b>1}, d['graph'])" || tail -5 gpurun_out/ab_$name.err
done
