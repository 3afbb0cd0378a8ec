#!/bin/bash
# A/B of other builds of the library (GA_LIB): bucketed parity tests under the first one, then C4 device steps for
# the default build and each alternative.  usage: gpu_altlib.sh alt1.so [alt2.so ...]
set -u
mkdir -p gpurun_out
GA_LIB=$1 timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider -k "scatter_kernels or bucketed or full_size_configs or segments or host_buffer" > gpurun_out/pytest_alt.log 2>&1
echo "pytest ($1) exit $?"; tail -4 gpurun_out/pytest_alt.log
specs="base"
for lib in "$@"; do n=$(basename $lib .so); specs="$specs ${n#libga_b200_}:GA_LIB=$lib"; done
bash scripts/gpu_ab_short.sh $specs
