#!/bin/bash
# gpurun -- bash scripts/gpu_ncu.sh <workload> <kernel-regex> <tag> [count] [skip]
W=${1:-c5}; K=${2:-sweep}; TAG=${3:-prof}; CNT=${4:-4}; SKIP=${5:-6}
mkdir -p gpurun_out
timeout 600 python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$TAG.json 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT -f -o gpurun_out/$TAG \
    python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu exit $?"; tail -c 300 gpurun_out/plain_$TAG.json; ls -la gpurun_out/$TAG.ncu-rep
