#!/bin/bash
# round 2, call H: wide-palette kernel v2 (dense walk queue, batched bitmaps) + warp/CTA-cooperative tail cutting: tests, config-4 palettes, full c4 bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "wide or narrow or tailcut or refgpu" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -8 gpurun_out/r2h_pytest.log
rm -f gpurun_out/r2h_quick.jsonl
run() { timeout 600 python bench.py --quick --steps 3 --warmup 1 "$@" >> gpurun_out/r2h_quick.jsonl 2>> gpurun_out/r2h_quick.err; }
for nc in 1024 2048 4096; do run --workload c4 --ncol $nc --traj 10; done
python - <<'PY'
import json
for l in open('gpurun_out/r2h_quick.jsonl'):
    d=json.loads(l); print(d['workload'], d['nCol'], d['kernel_mode'], round(d['ms_per_step'],3), round(d['chain_ms_per_sweep'],3), round(d['create_ms']), round(d['frac'],3), d['maxDeg'], d['traj'])
PY
for nc in 1024 2048; do timeout 900 python bench.py --workload c4 --ncol $nc --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench_c4_$nc.json 2> gpurun_out/r2h_bench_c4_$nc.err; tail -c 900 gpurun_out/r2h_bench_c4_$nc.json; tail -3 gpurun_out/r2h_bench_c4_$nc.err; done
