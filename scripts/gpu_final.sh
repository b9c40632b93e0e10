#!/bin/bash
# round-end evidence: tests, smoke, headline bench, ncu launch list + full capture of both sweep kernels (config 3)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench exit $?"; tail -c 2500 gpurun_out/bench_c3.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; tail -c 900 gpurun_out/bench_ref.json
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.json 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mcmcb200 -c 60 --csv --log-file gpurun_out/launches_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:blocked -s 6 -c 2 -f -o gpurun_out/final_c3 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/*.ncu-rep
