#!/bin/bash
# Round-end evidence on one B200: GPU parity tests, smoke, the default bench invocation, the reference arm, configs 5 and 4.
mkdir -p gpurun_out
TAG=${1:-r01i}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x -rs --no-header -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log; tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log; tail -2 gpurun_out/${TAG}_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; echo "bench default exit $?"; grep real gpurun_out/${TAG}_bench_default.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference arm exit $?"
for w in c5 c2; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; echo "bench $w exit $?"
done
python - <<PY
import json
for w in ("default","reference","c5","c2"):
    try:
        d=json.load(open("gpurun_out/${TAG}_bench_%s.json" % w))
        print(w, "ms %.4f value %.4g frac %s e2e %.4g ttc %s mode %s" % (d["ms_per_step"], d["value"], d.get("roofline",{}).get("frac"), d["e2e"]["value"], d.get("time_to_proper_coloring"), d.get("roofline",{}).get("kernel_mode")))
    except Exception as e: print(w, "failed", e)
PY
