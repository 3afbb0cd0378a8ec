#!/bin/bash
# One short visit: the tests added in this session (reference-made golden vectors of the bench's input class through the
# bucketed kernels, graph properties on the subsample and at full C4 size), then the user-API legs of C2 / C3 with the
# piecewise contig traversal.
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_c4_parity.py -q -m gpu -x --durations=8 \
    -k "mix or bench_input_class or properties or nodes_materialise or full_size_configs" > gpurun_out/new_tests.log 2>&1
echo "tests exit $?"; tail -14 gpurun_out/new_tests.log
for w in c3 c2; do
  timeout 200 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_${w}_trav.json 2> gpurun_out/bench_${w}_trav.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_${w}_trav.json').read().strip().splitlines()[-1]); print('$w', round(d['ms_per_step'],2), d['e2e'].get('user_api'))" || tail -5 gpurun_out/bench_${w}_trav.err
done
