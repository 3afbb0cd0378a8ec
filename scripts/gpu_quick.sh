#!/bin/bash
# quick GPU visit: bucketed-path parity tests, then C4 (and optionally C2) bench lines
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider -k "bucketed or full_size_configs" > gpurun_out/pytest_sk.log 2>&1
echo "pytest bucketed exit $?" | tee -a gpurun_out/pytest_sk.log
tail -8 gpurun_out/pytest_sk.log
for wl in ${WORKLOADS:-c4 c2}; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 2 > gpurun_out/bench_${wl}_q.json 2> gpurun_out/bench_${wl}_q.err
  echo "bench $wl exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${wl}_q.json"))
    print(d["ms_per_step"], d["value"]/1e9, d["roofline"]["kernel_ms_per_step"], d["roofline"]["whole_path"]["frac"], d["e2e"], d["graph"])
except Exception as e:
    print("no json", e)
PY
  tail -5 gpurun_out/bench_${wl}_q.err
done
