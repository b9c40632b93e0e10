#!/bin/bash
# gpurun -- bash scripts/gpu_c4.sh : BASELINE config 4 (R-MAT 50M): bench line, then ncu counters of the sweep launches
# (warp execution efficiency, DRAM bytes, sectors/request).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
for w in c4small c4; do
  timeout 1200 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w exit $?"; tail -c 2600 gpurun_out/bench_$w.json; tail -5 gpurun_out/bench_$w.err
done
M=smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed
timeout 1500 ncu --metrics $M --clock-control none -k regex:'sweep_kernel|blocked|binned' -s 3 -c 3 --csv --log-file gpurun_out/c4_counters.csv \
    python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1
echo "ncu exit $?"; tail -30 gpurun_out/c4_counters.csv | cut -c1-300
