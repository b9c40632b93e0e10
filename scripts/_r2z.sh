mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs --no-header -p no:cacheprovider > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu.log; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r02_smoke.log; tail -2 gpurun_out/r02_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; echo "bench default exit $?"; grep real gpurun_out/r02_bench_c3.err
timeout 900 python bench.py --workload c3 --proposal dynamic --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c3_dynamic.json 2> gpurun_out/r02_bench_c3_dynamic.err
timeout 900 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err
timeout 900 python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err
MCMCB200_LIB=$PWD/variants/libCheck.so python scripts/sanitize_driver.py > gpurun_out/r02_check_build_driver.log 2>&1; echo "check driver rc=$?" >> gpurun_out/r02_check_build_driver.log; tail -1 gpurun_out/r02_check_build_driver.log
MCMCB200_TEST_ANY_MODE=1 MCMCB200_LIB=$PWD/variants/libCheck.so timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/r02_check_build_pytest.log 2>&1; echo "pytest(check build) rc=$?" >> gpurun_out/r02_check_build_pytest.log; tail -2 gpurun_out/r02_check_build_pytest.log
python - <<'PY'
import json
for w in ("c3","c3_dynamic","c5","c2"):
    d=json.load(open("gpurun_out/r02_bench_%s.json" % w)); t=d.get("time_to_proper_coloring") or {}
    print(w, "ms %.4f frac %.4f chain %.4f e2e ms %.2f | ttc %.2f ms (%s sweeps, proper %s, setup %.1f) one-shot total %.1f" % (d["ms_per_step"], d["roofline"]["frac"], d["chain_ms_per_sweep"], d["e2e"]["ms_per_step"], t.get("ms"), t.get("sweeps"), t.get("proper"), t.get("setup_ms"), (t.get("one_shot") or {}).get("total_ms_with_setup")))
PY
