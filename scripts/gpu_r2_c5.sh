#!/bin/bash
# BASELINE config C5 on one B200: k sweep on the E. coli-sized read pairs of C3 (64-bit keys up to k = 32, 128-bit keys
# beyond) and the sketch width / row sweep of the -c route over the reference's three prime tables.
set -u
mkdir -p gpurun_out
out=gpurun_out/c5_sweep.jsonl; : > $out
for k in 21 25 29 31 33 41 51 63; do
  GA_BENCH_SKIP_E2E=1 timeout 600 python bench.py --workload c3 --k $k --steps 5 --warmup 3 --sample-reads 3000 2>/dev/null | tail -1 >> $out
  echo "k=$k exit $?"
done
for spec in "6e7 10" "1e7 10" "5e6 10" "1e7 8"; do
  set -- $spec
  GA_BENCH_SKIP_E2E=1 timeout 600 python bench.py --workload c3 --sketch --sketch-widths $1 --sketch-rows $2 --steps 5 --warmup 3 --sample-reads 3000 2>/dev/null | tail -1 >> $out
  echo "sketch widths=$1 rows=$2 exit $?"
done
python - <<'PY'
import json
for line in open("gpurun_out/c5_sweep.jsonl"):
    d = json.loads(line)
    c = d["config"]
    print("k=%-3d sketch=%-5s %7.2f ms/step  %6.2f G k-mers/s  whole-path %.3f  dominant %-22s cpu %.2f M/s  nodes %s" %
          (c["k"], c["sketch"], d["ms_per_step"], d["value"] / 1e9, d["roofline"]["whole_path"]["frac"], d["roofline"]["kernel"],
           d["cpu_baseline"]["value"] / 1e6, d["graph"]))
PY
