#!/bin/bash
# A/B of a second build of the library (GA_LIB): bucketed parity tests under it, then C4 device steps for both
set -u
mkdir -p gpurun_out
ALT=$1
GA_LIB=$ALT timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider -k "scatter_kernels or bucketed or full_size_configs or segments or host_buffer" > gpurun_out/pytest_alt.log 2>&1
echo "pytest (alt lib) exit $?"; tail -4 gpurun_out/pytest_alt.log
bash scripts/gpu_ab_short.sh base alt:GA_LIB=$ALT base2 alt2:GA_LIB=$ALT
