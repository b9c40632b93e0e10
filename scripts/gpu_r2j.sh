#!/bin/bash
# round 2, call J: wide kernel v3 (word-wise bitmap walk) tests + config 4 palettes + full c4 bench lines; config-3 register-budget variants; sanitizers
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "wide or narrow or tailcut or refgpu" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -4 gpurun_out/r2j_pytest.log
rm -f gpurun_out/r2j_quick.jsonl
run() { timeout 600 python bench.py --quick --steps 3 --warmup 1 "$@" >> gpurun_out/r2j_quick.jsonl 2>> gpurun_out/r2j_quick.err; }
runl() { lib=$1; shift; MCMCB200_LIB=$lib timeout 300 python bench.py --workload c3 --quick --steps 5 --warmup 3 "$@" >> gpurun_out/r2j_quick.jsonl 2>> gpurun_out/r2j_quick.err; }
for nc in 1024 2048 4096; do run --workload c4 --ncol $nc --traj 10; done
run --workload c4small --ncol 1024
V=$PWD/variants
for l in $PWD/mcmc_colorer_b200/libmcmcb200.so $V/libR56KU16.so $V/libR56PF4KU16.so $V/libR48PF4KU20.so; do runl $l; done
python - <<'PY'
import json
for l in open('gpurun_out/r2j_quick.jsonl'):
    d=json.loads(l); print(d['lib'].split('/')[-1], d['workload'], d['nCol'], d['kernel_mode'], round(d['ms_per_step'],3), round(d['chain_ms_per_sweep'],3), round(d['create_ms']), round(d['frac'],3), d['traj'])
PY
for nc in 1024 2048; do timeout 900 python bench.py --workload c4 --ncol $nc --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_c4_$nc.json 2> gpurun_out/r2j_bench_c4_$nc.err; tail -c 700 gpurun_out/r2j_bench_c4_$nc.json; tail -3 gpurun_out/r2j_bench_c4_$nc.err; done
bash scripts/gpu_sanitize.sh
