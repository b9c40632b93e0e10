#!/bin/bash
mkdir -p gpurun_out
M=smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct
timeout 900 ncu --metrics $M --clock-control none -k regex:'binned' -s 3 -c 1 --csv --log-file gpurun_out/c4_counters.csv \
    python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1
echo "ncu exit $?"
python - <<PY
import csv
for r in csv.reader(open("gpurun_out/c4_counters.csv")):
    if len(r)>10 and r[0].isdigit(): print(r[4][:40], r[-3], r[-2], r[-1])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:binned -s 3 -c 1 -f -o gpurun_out/r01h_c4small python bench.py --workload c4small --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c4s.log 2>&1; echo "full exit $?"
