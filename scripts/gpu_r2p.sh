#!/bin/bash
# round 2, call P: pass A split into a full-occupancy first part + the co-resident rest; finer first/last parts on small partitions
mkdir -p gpurun_out
rm -f gpurun_out/r2p_quick.jsonl gpurun_out/r2p_rank.jsonl
V=$PWD/variants; D=$PWD/mcmc_colorer_b200/libmcmcb200.so
for l in $D $V/libSplitA.so $V/libParts2.so $V/libSplitAParts2.so; do
  MCMCB200_LIB=$l timeout 300 python bench.py --workload c3 --quick --steps 5 --warmup 3 >> gpurun_out/r2p_quick.jsonl 2>> gpurun_out/r2p_quick.err
  echo "{\"lib\": \"$l\"}" >> gpurun_out/r2p_rank.jsonl
  MCMCB200_LIB=$l python scripts/rank_emulation.py 8 0 c3 >> gpurun_out/r2p_rank.jsonl 2>> gpurun_out/r2p_rank.err
done
MCMCB200_LIB=$V/libSplitA.so timeout 300 python bench.py --workload c5 --quick --steps 5 --warmup 3 >> gpurun_out/r2p_quick.jsonl 2>> gpurun_out/r2p_quick.err
python - <<'PY'
import json
for l in open('gpurun_out/r2p_quick.jsonl'):
    d=json.loads(l); print(d['lib'].split('/')[-1], d['workload'], d['kernel_mode'], round(d['ms_per_step'],3), round(d['chain_ms_per_sweep'],3), round(d['frac'],3), d['after_10'])
for l in open('gpurun_out/r2p_rank.jsonl'):
    d=json.loads(l); print(d.get('lib','').split('/')[-1] or (round(d['first_sweep_ms'],4), round(d['chain_ms_per_sweep'],4), d['kernel_mode']))
PY
MCMCB200_LIB=$V/libSplitA.so timeout 600 python -m pytest tests -m gpu -q -x -k "large_graph or free_running or config2" > gpurun_out/r2p_pytest.log 2>&1; tail -2 gpurun_out/r2p_pytest.log
