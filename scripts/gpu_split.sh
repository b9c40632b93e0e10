#!/bin/bash
# gpurun -- bash scripts/gpu_split.sh [lib] [cap] : per-kernel time and DRAM bytes of one sweep (ncu, few metrics)
mkdir -p gpurun_out
LIB=${1:-}; CAP=${2:-65504}
[ -n "$LIB" ] && export MCMCB200_LIB=$PWD/variants/lib_$LIB.so
export MCMCB200_STAGE_CAP_BYTES=$CAP
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum
timeout 900 ncu --metrics $M --clock-control none -k regex:'blocked' -s 6 -c 2 --csv --log-file gpurun_out/split_$LIB.csv \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_split.log 2>&1
echo "ncu exit $?"
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/split_$LIB.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows: print(r[4][:40], r[-3], r[-2], r[-1])
PY
