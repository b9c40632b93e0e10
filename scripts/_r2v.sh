mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2v_pytest.log
python scripts/create_bench.py c3 4 2>> gpurun_out/r2v.err | tee gpurun_out/r2v_create.jsonl | cut -c1-700
python scripts/create_bench.py c5 3 2>> gpurun_out/r2v.err | tee -a gpurun_out/r2v_create.jsonl | cut -c1-500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:blk_ --csv --log-file gpurun_out/r2v_build_launches.csv python bench.py --workload c3 --quick --steps 1 --warmup 1 > gpurun_out/r2v_ncu.log 2>&1
python - <<'PY'
import csv
from collections import defaultdict
rows=[r for r in csv.reader(open('gpurun_out/r2v_build_launches.csv')) if len(r)>10 and r[0].isdigit()]
agg=defaultdict(lambda:[0,0.0])
for r in rows:
    name=r[4].split('(')[0][:60]; v=float(r[-1].replace(',','')); u=r[-2]
    v = v/1000 if u=='us' else v/1e6 if u=='ns' else v*1000 if u=='s' else v
    agg[name][0]+=1; agg[name][1]+=v
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:6]: print(f"{t:9.3f} ms {c:3d} {k}")
PY
