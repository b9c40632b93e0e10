mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r02_smoke.log; tail -2 gpurun_out/r02_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; echo "bench default exit $?"; grep real gpurun_out/r02_bench_c3.err
timeout 900 python bench.py --workload c4heavy --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c4heavy.json 2> gpurun_out/r02_bench_c4heavy.err; echo "bench c4heavy exit $?"
timeout 900 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err; echo "bench c5 exit $?"
python scripts/create_bench.py c3 8 2>> gpurun_out/r2w.err > gpurun_out/r02_create_times.jsonl; python scripts/create_bench.py c5 6 2>> gpurun_out/r2w.err >> gpurun_out/r02_create_times.jsonl
python - <<'PY'
import json
for w in ("c3","c4heavy","c5"):
    d=json.load(open("gpurun_out/r02_bench_%s.json" % w)); t=d.get("time_to_proper_coloring") or {}
    print(w, "ms %.4f frac %.4f e2e ms %.2f | ttc %.2f ms (%s sweeps, proper %s, setup %.1f) one-shot total %.1f" % (d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], t.get("ms"), t.get("sweeps"), t.get("proper"), t.get("setup_ms"), (t.get("one_shot") or {}).get("total_ms_with_setup")))
for l in open('gpurun_out/r02_create_times.jsonl'):
    d=json.loads(l); print(d['workload'], [c['create_ms'] for c in d['creates']])
PY
