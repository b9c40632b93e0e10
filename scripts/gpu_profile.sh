#!/bin/bash
# Under gpurun: headline bench (plain), then the ncu launch list and one full capture of the sweep kernel on a
# smaller workload (ncu replays each kernel ~40x).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
W=${1:-c5}
timeout 1500 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 exit $?"; tail -c 3000 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$W.json 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$W.csv \
    python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$W.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 2 -f -o gpurun_out/prof_$W \
    python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$W.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/
