#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 1500 python scripts/quality_sweep.py --n 1000000 --big 10000000 --seeds 3 > gpurun_out/quality.log 2>&1; echo "quality exit $?"; tail -3 gpurun_out/quality.log | cut -c1-400
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.json 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"blocked_|class_sizes|init_colors|narrow_|widen_|finalize" -c 80 --csv --log-file gpurun_out/launches_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
