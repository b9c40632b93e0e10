#!/bin/bash
# round 2, call M (8 GPUs): one rank alone (no peers) against the 8-GPU step with the device-side start barrier
N=${1:-8}
mkdir -p gpurun_out
python scripts/rank_emulation.py $N 0 c3 > gpurun_out/r2m_rank0_of_$N.json 2> gpurun_out/r2m_rank.err; cat gpurun_out/r2m_rank0_of_$N.json
python scripts/rank_emulation.py $N 3 c3 > gpurun_out/r2m_rank3_of_$N.json 2>> gpurun_out/r2m_rank.err; cat gpurun_out/r2m_rank3_of_$N.json
for W in c3; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --workload $W --steps 10 --warmup 3 > gpurun_out/r2m_bench_${W}_x$N.json 2> gpurun_out/r2m_bench_${W}_x$N.err; echo "bench $W exit $?"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2m_bench_${W}_x$N.json'))
    print('$W', d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'chain', round(d['chain_ms_per_sweep'],4), 'frac', round(d['roofline']['frac'],3), d['roofline']['kernel'], 'e2e ms', round(d['e2e']['ms_per_step'],2), d['after_10_chain_sweeps'])
except Exception as e: print('no json', e)
PY
tail -2 gpurun_out/r2m_bench_${W}_x$N.err
done
MCMCB200_NO_P2P=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload c3 --steps 10 --warmup 3 > gpurun_out/r2m_bench_c3_nccl_x$N.json 2> gpurun_out/r2m_bench_c3_nccl_x$N.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2m_bench_c3_nccl_x$N.json'))
    print('c3 NCCL path', d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'chain', round(d['chain_ms_per_sweep'],4))
except Exception as e: print('no json', e)
PY
