#!/bin/bash
# gpurun -- bash scripts/gpu_splits.sh lib ... : ncu per-kernel time + DRAM bytes for several variants (pass A/B of one sweep)
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for LIB in "$@"; do
  MCMCB200_LIB=$PWD/variants/lib_$LIB.so timeout 600 ncu --metrics $M --clock-control none -k regex:'blocked' -s 6 -c 2 --csv --log-file gpurun_out/split_$LIB.csv \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_split.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/split_$LIB.csv")) if len(r)>10 and r[0].isdigit()]
out={}
for r in rows: out.setdefault(r[4][5:20],{})[r[-3].split('.')[0][-12:]]=float(r[-1])
for k,v in out.items(): print("$LIB", k, " ".join(f"{a}={b/1e9 if b>1e7 else b/1e6:.3f}" for a,b in v.items()))
PY
done
