#!/bin/bash
# Round-2 check on one B200: bucket-path parity tests, then the C4 bench on this build and (A/B) on the
# round-1 library kept under genome-assembler_b200/build/libga_b200_r1.so.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_c4_parity.py tests/test_gpu_parity.py -q -m gpu -x --durations=10 \
    -k "${GA_TEST_FILTER:-c4 or bucket or segments or host_buffer or determin or sweep or prefilter or scatter}" \
    > gpurun_out/r2_check_tests.log 2>&1
echo "tests exit $?"; tail -25 gpurun_out/r2_check_tests.log
run() {
  name=$1; shift
  ( for e in "$@"; do export "$e"; done
    GA_BENCH_SKIP_E2E=1 timeout 600 python bench.py --workload c4 --steps 3 --warmup 2 --sample-reads 2000 \
      > gpurun_out/bench_r2_$name.json 2> gpurun_out/bench_r2_$name.err )
  python -c "
import json; d=json.load(open('gpurun_out/bench_r2_$name.json')); k=d['roofline']['kernel_ms_per_step']; print('$name', round(d['ms_per_step'],1), {a:round(b,1) for a,b in k.items() if b>1}, d['graph'])" || tail -5 gpurun_out/bench_r2_$name.err
}
run new GA_TRACE=1
run r1 GA_LIB=$PWD/genome-assembler_b200/build/libga_b200_r1.so GA_SK_SLOTS=8192
run new2
