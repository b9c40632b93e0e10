#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -rs --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for w in c4small c4; do
  timeout 1200 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json')); print('$w', d['ms_per_step'], d['chain_ms_per_sweep'], d['roofline']['frac'], d['edges_per_sec'], d['roofline'].get('kernel_mode'))"; tail -3 gpurun_out/bench_$w.err
done
