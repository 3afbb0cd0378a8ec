#!/bin/bash
# multi-GPU visit (run with gpurun --gpus N): parity of the sharded build, then bench at N and at 1
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1
echo "pytest multi exit $?"; tail -15 gpurun_out/pytest_multi.log
for wl in ${WORKLOADS:-c4}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29733 \
     bench.py --gpus $N --workload $wl --steps 3 --warmup 2 > gpurun_out/bench_${wl}_n$N.json 2> gpurun_out/bench_${wl}_n$N.err
  echo "bench $wl n=$N exit $?"; tail -3 gpurun_out/bench_${wl}_n$N.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_${wl}_n$N.json") if l.startswith("{")][-1])
    print(d["n_gpus"], d["ms_per_step"], d["value"]/1e9, d["roofline"]["kernel_ms_per_step"], d["e2e"], d["graph"])
except Exception as e:
    print("no json", e)
PY
done
