#!/bin/bash
# Closing visit of round 2 on one B200 (short on GPU minutes: no ncu here, the kernels have not changed since
# profiles/r02/*_final.*): the whole GPU suite on the final build, smoke(), the default bench line (C4), the C2 / C3
# lines with the user-API leg, and the host-side profile of the user API on C3.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x --durations=12 > gpurun_out/pytest_gpu_close.log 2>&1
echo "tests exit $?"; tail -18 gpurun_out/pytest_gpu_close.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_close.log 2>&1
echo "smoke exit $?"; tail -3 gpurun_out/smoke_close.log
timeout 300 python bench.py > gpurun_out/bench_c4_close.json 2> gpurun_out/bench_c4_close.err
echo "bench exit $?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_c4_close.json').read().strip().splitlines()[-1]); r = d['roofline']
print(round(d['ms_per_step'], 1), 'ms', round(d['value'] / 1e9, 1), 'G/s', {a: round(b, 1) for a, b in r['kernel_ms_per_step'].items()})
print('frac', round(r['frac'], 3), 'whole', round(r['whole_path']['frac'], 3), 'e2e', round(d['e2e']['ms_per_step'], 1), d['clocks'], d['gpu_launches'])
PY
for w in c2 c3; do
  timeout 200 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_${w}_close.json 2> gpurun_out/bench_${w}_close.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_${w}_close.json').read().strip().splitlines()[-1]); print('$w', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {a: round(b, 2) for a, b in d['roofline']['kernel_ms_per_step'].items()}, d['e2e'].get('user_api'))" || tail -5 gpurun_out/bench_${w}_close.err
done
timeout 100 python scripts/profile_user_api.py c3 3 > gpurun_out/prof_user_c3.txt 2>&1
grep -n "trace\|one traced" gpurun_out/prof_user_c3.txt | tail -30
sed -n '/Ordered by: internal time/,$p' gpurun_out/prof_user_c3.txt | head -34
