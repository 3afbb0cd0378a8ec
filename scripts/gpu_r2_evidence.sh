#!/bin/bash
# Round-2 evidence on one B200: GPU suite, default bench line (C4), ncu launch list of the same command, DRAM bytes of
# the bucketed kernels, a full-set capture on a quarter-size instance, then the small configurations and the reference arm.
# Every ncu pass runs only after the same command has exited 0 without ncu.
set -u
mkdir -p gpurun_out
if [ -z "${GA_SKIP_TESTS:-}" ]; then
  timeout 1500 python -m pytest tests -q -m gpu -x --durations=15 > gpurun_out/pytest_gpu.log 2>&1
  echo "tests exit $?"; tail -22 gpurun_out/pytest_gpu.log
fi
timeout 900 python bench.py > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
echo "bench exit $?"; python - <<'EOF'
import json
d = json.load(open('gpurun_out/bench_c4.json')); r = d['roofline']
print(round(d['ms_per_step'], 1), 'ms', round(d['value'] / 1e9, 1), 'G/s', {a: round(b, 1) for a, b in r['kernel_ms_per_step'].items()})
print('frac', round(r['frac'], 3), 'whole', round(r['whole_path']['frac'], 3), 'e2e', d['e2e'], d['clocks'], d['gpu_launches'], d['memory_gb'])
EOF
CMD="python bench.py --workload c4 --steps 2 --warmup 1 --sample-reads 2000"
GA_BENCH_SKIP_E2E=1 $CMD > gpurun_out/plain_c4.json 2> gpurun_out/plain_c4.err
echo "plain exit $?"
GA_BENCH_SKIP_E2E=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_c4.csv $CMD > gpurun_out/ncu_launch_c4.log 2>&1
echo "launch list exit $?"
GA_BENCH_SKIP_E2E=1 timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'sk_' -s 7 -c 7 --csv \
    --log-file gpurun_out/dram_c4.csv $CMD > gpurun_out/ncu_dram_c4.log 2>&1
echo "dram exit $?"
R=25000000
Q="python bench.py --workload c4 --reads $R --genome $((R / 2)) --steps 1 --warmup 1 --sample-reads 2000"
GA_BENCH_SKIP_E2E=1 $Q > gpurun_out/plain_quarter.json 2> gpurun_out/plain_quarter.err
echo "quarter plain exit $?"
GA_BENCH_SKIP_E2E=1 timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:'sk_' -s 7 -c 7 \
    -o gpurun_out/prof_sk -f $Q > gpurun_out/ncu_full_sk.log 2>&1
echo "full capture exit $?"
ncu -i gpurun_out/prof_sk.ncu-rep --page raw --csv > gpurun_out/prof_sk_raw.csv 2>/dev/null
ls -la gpurun_out/prof_sk.ncu-rep gpurun_out/prof_sk_raw.csv
if [ -z "${GA_SKIP_SMALL:-}" ]; then
  bash scripts/gpu_r2_small.sh
  timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
  echo "reference exit $?"; cat gpurun_out/bench_reference.json
fi
