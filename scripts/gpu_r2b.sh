#!/bin/bash
# round 2, call B: GPU test suite + smoke, then geometry / stage-size sweep of the blocked sweep on config 3
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2b_smoke.log
tail -2 gpurun_out/r2b_smoke.log
run() { lib=$1; shift; MCMCB200_LIB=$lib timeout 300 python bench.py --workload c3 --quick --steps 5 --warmup 3 "$@" >> gpurun_out/r2b_quick.jsonl 2>> gpurun_out/r2b_quick.err; }
D=$PWD/mcmc_colorer_b200/libmcmcb200.so
for s in 24576 28672 32768 36864 40960; do run $D --stage-cap-bytes $s; done
for s in 24576 28672 32768; do run $PWD/variants/libB256.so --stage-cap-bytes $s; done
for s in 32768 40960; do run $PWD/variants/libB320.so --stage-cap-bytes $s; done
run $PWD/variants/libA128.so --stage-cap-bytes 32768
for s in 32768 45056; do run $PWD/variants/libA128B448.so --stage-cap-bytes $s; done
run $PWD/variants/libA384B320.so --stage-cap-bytes 32768
run $PWD/variants/libA512B256.so --stage-cap-bytes 28672
run $D --stage-cap-bytes 32768 --item-bits 17
run $D --stage-cap-bytes 32768 --item-bits 16
python - <<'PY'
import json
for l in open('gpurun_out/r2b_quick.jsonl'):
    d=json.loads(l); print(d['lib'].split('/')[-1], d['tuning'], d['kernel_mode'], round(d['ms_per_step'],3), round(d['chain_ms_per_sweep'],3), round(d['create_ms']))
PY
