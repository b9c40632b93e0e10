#!/bin/bash
# round 2, call D: GPU tests, prefetch A/B, then ONE ncu --set full capture of the two blocked kernels on config 3
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -15 gpurun_out/r2d_pytest.log
rm -f gpurun_out/r2d_quick.jsonl
run() { lib=$1; shift; MCMCB200_LIB=$lib timeout 300 python bench.py --workload c3 --quick --steps 5 --warmup 3 "$@" >> gpurun_out/r2d_quick.jsonl 2>> gpurun_out/r2d_quick.err; }
D=$PWD/mcmc_colorer_b200/libmcmcb200.so
run $D
run $PWD/variants/libNoPf.so
run $D --proposal dynamic
run $D --workload c5
run $D --workload c2
python - <<'PY'
import json
for l in open('gpurun_out/r2d_quick.jsonl'):
    d=json.loads(l); print(d['lib'].split('/')[-1], d['workload'], d['tuning'], d['kernel_mode'], round(d['ms_per_step'],3), round(d['chain_ms_per_sweep'],3), round(d['create_ms']), round(d['frac'],3))
PY
python bench.py --workload c3 --quick --steps 2 --warmup 3 > gpurun_out/r2d_plain.json 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:blocked -s 6 -c 2 -o gpurun_out/r2d_c3 python bench.py --workload c3 --quick --steps 2 --warmup 3 > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_ncu.log
ls -la gpurun_out/*.ncu-rep
