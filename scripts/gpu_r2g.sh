#!/bin/bash
# round 2, call G: wide-palette kernel -- GPU tests, then the violation trajectory of config 4 (small and full) at several palettes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "wide or narrow or tailcut" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -15 gpurun_out/r2g_pytest.log
rm -f gpurun_out/r2g_quick.jsonl
run() { timeout 600 python bench.py --quick --steps 3 --warmup 1 "$@" >> gpurun_out/r2g_quick.jsonl 2>> gpurun_out/r2g_quick.err; }
for nc in 512 1024 4096; do run --workload c4small --ncol $nc --traj 80; done
for nc in 512 1024 2048 4096 16384; do run --workload c4 --ncol $nc --traj 40; done
python - <<'PY'
import json
for l in open('gpurun_out/r2g_quick.jsonl'):
    d=json.loads(l); print(d['workload'], d['nCol'], d['kernel_mode'], round(d['ms_per_step'],3), round(d['chain_ms_per_sweep'],3), round(d['create_ms']), round(d['frac'],3), d['maxDeg'], d['traj'])
PY
tail -5 gpurun_out/r2g_quick.err
