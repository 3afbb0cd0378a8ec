#!/bin/bash
# Round-2 multi-GPU visit (gpurun --gpus N): bit-identity of the sharded build (both exchanges), then the C4 bench at N
# with the NVLink push exchange and, for comparison, with the round-1 NCCL exchange.
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 \
    scripts/multi_check.py > gpurun_out/multi_check_n$N.log 2>&1
echo "multi_check exit $?"; grep -E "multi_check|Error|error" gpurun_out/multi_check_n$N.log | tail -12
for ex in ${EXCHANGES:-push nccl}; do
  GA_MULTI_EXCHANGE=$ex timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
     --master-port 29733 bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench_c4_n${N}_$ex.json 2> gpurun_out/bench_c4_n${N}_$ex.err
  echo "bench n=$N $ex exit $?"; tail -3 gpurun_out/bench_c4_n${N}_$ex.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_c4_n${N}_$ex.json") if l.startswith("{")][-1])
    r=d["roofline"]
    print(d["n_gpus"], round(d["ms_per_step"],1), round(d["value"]/1e9,1), {a:round(b,1) for a,b in r["kernel_ms_per_step"].items()})
    print("  stages", {a:round(b,1) for a,b in r["stage_ms_per_step"].items()})
    print("  e2e", d["e2e"], d["graph"])
except Exception as e:
    print("no json", e)
PY
done
