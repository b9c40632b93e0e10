#!/bin/bash
# round 2, call E: GPU tests of the early-ticket pass B, then pass-A bytes-in-flight (A_KU) x early-ticket matrix on config 3
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -5 gpurun_out/r2e_pytest.log
rm -f gpurun_out/r2e_quick.jsonl
run() { lib=$1; shift; MCMCB200_LIB=$lib timeout 300 python bench.py --workload c3 --quick --steps 5 --warmup 3 "$@" >> gpurun_out/r2e_quick.jsonl 2>> gpurun_out/r2e_quick.err; }
D=$PWD/mcmc_colorer_b200/libmcmcb200.so
for l in $D $PWD/variants/libNoET.so $PWD/variants/libKU12.so $PWD/variants/libKU12NoET.so $D; do run $l; done
run $D --workload c5
python - <<'PY'
import json
for l in open('gpurun_out/r2e_quick.jsonl'):
    d=json.loads(l); print(d['lib'].split('/')[-1], d['workload'], d['tuning'], d['kernel_mode'], round(d['ms_per_step'],3), round(d['chain_ms_per_sweep'],3), round(d['create_ms']), round(d['frac'],3))
PY
