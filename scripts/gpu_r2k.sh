#!/bin/bash
# round 2, call K: flattened warp-row batches in the wide kernel, new config 4 (Reddit-like skew, nCol = maxDeg), bounded tail cutting,
# expectedSweeps, the bounds-checking build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "wide or narrow or tailcut or refgpu or argument" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -4 gpurun_out/r2k_pytest.log
MCMCB200_LIB=$PWD/variants/libCheck.so python scripts/sanitize_driver.py > gpurun_out/r2k_check_driver.log 2>&1; echo "check driver rc=$?" >> gpurun_out/r2k_check_driver.log; tail -3 gpurun_out/r2k_check_driver.log
MCMCB200_LIB=$PWD/variants/libCheck.so timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2k_check_pytest.log 2>&1; echo "pytest(check build) rc=$?" >> gpurun_out/r2k_check_pytest.log; tail -3 gpurun_out/r2k_check_pytest.log
rm -f gpurun_out/r2k_quick.jsonl
run() { timeout 600 python bench.py --quick --steps 3 --warmup 1 "$@" >> gpurun_out/r2k_quick.jsonl 2>> gpurun_out/r2k_quick.err; }
run --workload c4 --traj 20
run --workload c4heavy --traj 10
run --workload c4heavy --ncol 512
run --workload c4small
run --workload c3 --expected-sweeps 8
run --workload c5 --stage-cap-bytes 32768
run --workload c5 --stage-cap-bytes 45056
run --workload c5
python - <<'PY'
import json
for l in open('gpurun_out/r2k_quick.jsonl'):
    d=json.loads(l); print(d['workload'], d['nCol'], d['maxDeg'], d['nnz'], d['kernel_mode'], round(d['ms_per_step'],3), round(d['chain_ms_per_sweep'],3), round(d['create_ms']), round(d['frac'],3), d['tuning'], d['traj'])
PY
tail -3 gpurun_out/r2k_quick.err
timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/r2k_bench_c4.json 2> gpurun_out/r2k_bench_c4.err; tail -c 1500 gpurun_out/r2k_bench_c4.json; tail -3 gpurun_out/r2k_bench_c4.err
