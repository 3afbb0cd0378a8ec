#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-2}
GA_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29733 \
     bench.py --gpus $N --workload c4 --steps 2 --warmup 1 --sample-reads 2000 > gpurun_out/trace_multi.log 2>&1
grep "trace r0" gpurun_out/trace_multi.log | tail -45
grep "^{" gpurun_out/trace_multi.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['e2e'])"
timeout 600 python bench.py --workload c4 --steps 3 --warmup 2 --sample-reads 2000 > gpurun_out/bench_c4_q.json 2>gpurun_out/bench_c4_q.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_q.json')); print(d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['e2e'])"
