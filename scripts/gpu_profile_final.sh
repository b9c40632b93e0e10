#!/bin/bash
# Round-end evidence: bench line (plain), ncu launch list of the same command, one --set full capture of pass A and pass B.
mkdir -p gpurun_out
TAG=${1:-r01g}
timeout 900 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_c3.csv \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "launch list exit $?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:blocked -s 6 -c 2 -f -o gpurun_out/${TAG}_c3 \
    python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/${TAG}*
