#!/bin/bash
# gpu tests on the default build, then the kernel variants
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -rs --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
bash scripts/gpu_variants.sh "$@"
