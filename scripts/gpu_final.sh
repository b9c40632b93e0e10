#!/bin/bash
# Round-end evidence on one B200: GPU parity tests, smoke, the default bench invocation, the reference arm, the other configs,
# the ncu launch list and one `ncu --set full` capture of the hot kernels (each after the same command ran clean without ncu).
#   gpurun --timeout 2400 -- bash scripts/gpu_final.sh r02
mkdir -p gpurun_out
TAG=${1:-r02}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rs --no-header -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log; tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log; tail -2 gpurun_out/${TAG}_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench default exit $?"; grep real gpurun_out/${TAG}_bench_c3.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference arm exit $?"
for w in c5 c2 c4 c4heavy; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; echo "bench $w exit $?"
done
timeout 900 python bench.py --workload c3 --proposal dynamic --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c3_dynamic.json 2> gpurun_out/${TAG}_bench_c3_dynamic.err; echo "bench c3 dynamic exit $?"
python - <<PY
import json
for w in ("c3","reference","c5","c2","c4","c4heavy","c3_dynamic"):
    try:
        d=json.load(open("gpurun_out/${TAG}_bench_%s.json" % w))
        t=d.get("time_to_proper_coloring") or {}
        print(w, "ms %.4f value %.4g frac %s e2e %.4g mode %s | ttc %s ms (%s sweeps, proper %s, setup %s) one-shot total %s" % (d["ms_per_step"], d["value"], d.get("roofline",{}).get("frac"), d["e2e"]["value"],
              d.get("roofline",{}).get("kernel_mode"), t.get("ms"), t.get("sweeps"), t.get("proper"), t.get("setup_ms"), (t.get("one_shot") or {}).get("total_ms_with_setup")))
    except Exception as e: print(w, "failed", e)
PY
# launch list + full captures (numbers printed under ncu are never bench values)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_c3.csv python bench.py --workload c3 --quick --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:blocked -s 6 -c 2 -f -o gpurun_out/${TAG}_c3_blocked python bench.py --workload c3 --quick --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_c3.log 2>&1; echo "ncu c3 exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:wide_ -s 4 -c 4 -f -o gpurun_out/${TAG}_c4_wide python bench.py --workload c4 --quick --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu_c4.log 2>&1; echo "ncu c4 exit $?"
ls -la gpurun_out/${TAG}*.ncu-rep
