#!/bin/bash
# lane-per-read scatter: record-level agreement with the warp kernel, bucketed parity tests, then A/B on C4
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider -k "scatter_kernels or bucketed or full_size_configs" > gpurun_out/pytest_lane.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/pytest_lane.log
bash scripts/gpu_ab_short.sh ${VARIANTS:-warp:GA_SK_SCATTER=warp lane:GA_SK_SCATTER=lane lane128:GA_SK_SCATTER=lane128}
