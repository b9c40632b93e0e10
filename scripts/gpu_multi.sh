#!/bin/bash
# gpurun --gpus N -- bash scripts/gpu_multi.sh N [workload]
N=${1:-2}; W=${2:-c5}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 scripts/multigpu_check.py > gpurun_out/multigpu_check_$N.log 2>&1; echo "check exit $?"; tail -5 gpurun_out/multigpu_check_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --workload $W --steps 5 --warmup 3 > gpurun_out/bench_${W}_x$N.json 2> gpurun_out/bench_${W}_x$N.err; echo "bench exit $?"; tail -c 2200 gpurun_out/bench_${W}_x$N.json; tail -4 gpurun_out/bench_${W}_x$N.err
