#!/bin/bash
# round 2, call N: pass-A item size for 8-GPU-sized partitions (one rank alone), wide kernel with deferred long rows
mkdir -p gpurun_out
python scripts/rank_emulation.py 8 0 c3 0,15,14,13 > gpurun_out/r2n_rank.jsonl 2> gpurun_out/r2n_rank.err; cat gpurun_out/r2n_rank.jsonl | cut -c1-400
python scripts/rank_emulation.py 8 0 c3 0,14 65504 >> gpurun_out/r2n_rank.jsonl 2>> gpurun_out/r2n_rank.err; tail -2 gpurun_out/r2n_rank.jsonl | cut -c1-400
python scripts/rank_emulation.py 4 0 c3 0,15,14 >> gpurun_out/r2n_rank.jsonl 2>> gpurun_out/r2n_rank.err; tail -3 gpurun_out/r2n_rank.jsonl | cut -c1-400
timeout 900 python -m pytest tests -m gpu -q -x -k "wide or tailcut" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log; tail -3 gpurun_out/r2n_pytest.log
rm -f gpurun_out/r2n_quick.jsonl
run() { timeout 600 python bench.py --quick --steps 3 --warmup 1 "$@" >> gpurun_out/r2n_quick.jsonl 2>> gpurun_out/r2n_quick.err; }
run --workload c4
run --workload c4heavy
run --workload c4small
python - <<'PY'
import json
for l in open('gpurun_out/r2n_quick.jsonl'):
    d=json.loads(l); print(d['workload'], d['nCol'], d['kernel_mode'], round(d['ms_per_step'],3), round(d['chain_ms_per_sweep'],3), round(d['frac'],3))
PY
