#!/bin/bash
# round 2, call C: full GPU test suite (incl. the reference-GPU-kernel parity tests) + tile-size / item-size matrix on config 3
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -25 gpurun_out/r2c_pytest.log
rm -f gpurun_out/r2c_quick.jsonl
run() { timeout 300 python bench.py --workload c3 --quick --steps 5 --warmup 3 "$@" >> gpurun_out/r2c_quick.jsonl 2>> gpurun_out/r2c_quick.err; }
for s in 17408 25600 32768 40960; do for b in 15 16 17; do run --stage-cap-bytes $s --item-bits $b; done; done
run --stage-cap-bytes 17408 --item-bits 16 --stage-buffers 2
run --stage-cap-bytes 32768 --item-bits 14
python - <<'PY'
import json
for l in open('gpurun_out/r2c_quick.jsonl'):
    d=json.loads(l); print(d['tuning'], d['kernel_mode'], round(d['ms_per_step'],3), round(d['chain_ms_per_sweep'],3), round(d['create_ms']))
PY
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --stage-cap-bytes 32768 --item-bits 16 > gpurun_out/r2c_bench_c3.json 2> gpurun_out/r2c_bench_c3.err; tail -c 1500 gpurun_out/r2c_bench_c3.json
