#!/bin/bash
# round 2, call F: pass-A bytes in flight (A_KU) x pass-B CTA width matrix on config 3 (early ticket off)
mkdir -p gpurun_out
rm -f gpurun_out/r2f_quick.jsonl
run() { lib=$1; shift; MCMCB200_LIB=$lib timeout 300 python bench.py --workload c3 --quick --steps 5 --warmup 3 "$@" >> gpurun_out/r2f_quick.jsonl 2>> gpurun_out/r2f_quick.err; }
D=$PWD/mcmc_colorer_b200/libmcmcb200.so
V=$PWD/variants
for rep in 1 2; do for l in $D $V/libKU10.so $V/libKU12.so $V/libB352KU16.so $V/libB320KU20.so $V/libB352KU12.so; do run $l; done; done
for l in $V/libKU12.so $V/libB352KU16.so $V/libB320KU20.so; do for s in 28672 36864 40960; do run $l --stage-cap-bytes $s; done; done
for l in $V/libKU12.so; do for b in 15 17; do run $l --item-bits $b; done; done
run $V/libKU12.so --workload c5
run $V/libKU12.so --workload c2
python - <<'PY'
import json
for l in open('gpurun_out/r2f_quick.jsonl'):
    d=json.loads(l); print(d['lib'].split('/')[-1], d['workload'], d['tuning'], d['kernel_mode'], round(d['ms_per_step'],3), round(d['chain_ms_per_sweep'],3), round(d['create_ms']), round(d['frac'],3))
PY
