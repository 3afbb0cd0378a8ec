#!/bin/bash
# N=1 tests + bench, then N-GPU parity + bench (run with gpurun --gpus N)
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider -k "bucketed or full_size_configs or multi" > gpurun_out/pytest_sk.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_sk.log
timeout 600 python bench.py --workload c4 --steps 3 --warmup 2 --sample-reads 2000 > gpurun_out/bench_c4_q.json 2>gpurun_out/bench_c4_q.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_q.json')); print(1, d['ms_per_step'], d['value']/1e9, d['roofline']['kernel_ms_per_step'], d['e2e'])"; tail -3 gpurun_out/bench_c4_q.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29733 \
     bench.py --gpus $N --workload c4 --steps 3 --warmup 2 --sample-reads 2000 > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err
grep "^{" gpurun_out/bench_c4_n$N.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['value']/1e9, d['roofline']['kernel_ms_per_step'], d['e2e'])"; tail -3 gpurun_out/bench_c4_n$N.err
