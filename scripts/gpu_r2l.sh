#!/bin/bash
# round 2, call L: two-phase wide sweep (count + walk kernels): tests, bounds-checking build over the whole parity suite, config 4 lines, ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "wide or narrow or tailcut or refgpu or argument or large_graph" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -4 gpurun_out/r2l_pytest.log
MCMCB200_LIB=$PWD/variants/libCheck.so python scripts/sanitize_driver.py > gpurun_out/r2l_check_driver.log 2>&1; echo "check driver rc=$?" >> gpurun_out/r2l_check_driver.log; tail -2 gpurun_out/r2l_check_driver.log
MCMCB200_TEST_ANY_MODE=1 MCMCB200_LIB=$PWD/variants/libCheck.so timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2l_check_pytest.log 2>&1; echo "pytest(check build) rc=$?" >> gpurun_out/r2l_check_pytest.log; tail -3 gpurun_out/r2l_check_pytest.log
rm -f gpurun_out/r2l_quick.jsonl
run() { timeout 600 python bench.py --quick --steps 3 --warmup 1 "$@" >> gpurun_out/r2l_quick.jsonl 2>> gpurun_out/r2l_quick.err; }
run --workload c4 --traj 10
run --workload c4heavy --traj 5
run --workload c4small
run --workload c4 --proposal dynamic
python - <<'PY'
import json
for l in open('gpurun_out/r2l_quick.jsonl'):
    d=json.loads(l); print(d['workload'], d['nCol'], d['maxDeg'], d['nnz'], d['kernel_mode'], round(d['ms_per_step'],3), round(d['chain_ms_per_sweep'],3), round(d['create_ms']), round(d['frac'],3), d['traj'])
PY
tail -3 gpurun_out/r2l_quick.err
ncu --set full --clock-control none --import-source on -k regex:wide_ -s 3 -c 3 -o gpurun_out/r2l_c4_wide python bench.py --workload c4 --quick --steps 2 --warmup 1 > gpurun_out/r2l_ncu.log 2>&1
ls -la gpurun_out/r2l*.ncu-rep
