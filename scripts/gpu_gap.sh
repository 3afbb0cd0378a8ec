#!/bin/bash
set -u
mkdir -p gpurun_out
python - <<'PY'
import torch, time
x=torch.empty(4<<30,dtype=torch.uint8,pin_memory=True); d=torch.empty(4<<30,dtype=torch.uint8,device='cuda')
for _ in range(2):
    torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(x,non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print('H2D GB/s', 4.29/dt)
    torch.cuda.synchronize(); t=time.perf_counter(); x.copy_(d,non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print('D2H GB/s', 4.29/dt)
PY
for i in 1 2 3; do
timeout 600 python bench.py --workload c4 --steps 3 --warmup 2 --sample-reads 2000 > gpurun_out/bench_gap.json 2>gpurun_out/bench_gap.err; python -c "
import json; d=json.load(open('gpurun_out/bench_gap.json')); k=d['roofline']['kernel_ms_per_step']; print(d['ms_per_step'], sum(k.values()), d['e2e']['ms_per_step']); print({a:round(b,2) for a,b in d['roofline']['stage_ms_per_step'].items()})"
done
GA_TRACE=1 timeout 600 python bench.py --workload c4 --steps 1 --warmup 1 --sample-reads 2000 2>&1 | grep trace | tail -18
