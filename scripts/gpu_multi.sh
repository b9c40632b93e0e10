#!/bin/bash
# gpurun --gpus N -- bash scripts/gpu_multi.sh N : bit-identity check, then the N-GPU bench lines of configs 3 and 4
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 scripts/multigpu_check.py > gpurun_out/multigpu_check_$N.log 2>&1; echo "check exit $?"; tail -3 gpurun_out/multigpu_check_$N.log
for W in c3 c4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --workload $W --steps 5 --warmup 3 > gpurun_out/bench_${W}_x$N.json 2> gpurun_out/bench_${W}_x$N.err; echo "bench $W exit $?"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${W}_x$N.json'))
    print('$W', d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), d['roofline']['kernel'], 'e2e ms', round(d['e2e']['ms_per_step'],2), d['after_10_chain_sweeps'], d['config']['parallelism'][:60])
except Exception as e: print('no json', e)
PY
tail -2 gpurun_out/bench_${W}_x$N.err
done
