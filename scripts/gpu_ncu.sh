#!/bin/bash
# gpurun -- bash scripts/gpu_ncu.sh <workload> <kernel-regex> <tag>
W=${1:-c5}; K=${2:-sweep}; TAG=${3:-prof}
mkdir -p gpurun_out
timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$TAG.json 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 4 -f -o gpurun_out/$TAG \
    python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu exit $?"; tail -c 600 gpurun_out/plain_$TAG.json; ls -la gpurun_out/$TAG.ncu-rep
