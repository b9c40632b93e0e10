#!/bin/bash
# quick kernel experiments: rebuild with EXTRA flags on the box, run the c5/c3 sweeps with different stage caps
mkdir -p gpurun_out
for T in 256 512; do
  make -s -C mcmc_colorer_b200/csrc clean; make -s -C mcmc_colorer_b200/csrc EXTRA="-DMCMCB200_THREADS_B=$T" > /dev/null 2>&1
  for CAP in 24576 32768 49152 65504; do
    for W in c5 c3; do
      MCMCB200_STAGE_CAP_BYTES=$CAP timeout 300 python bench.py --workload $W --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('threadsB $T cap $CAP $W ms %.3f frac %.3f chain %.3f' % (d['ms_per_step'], d['roofline']['frac'], d['chain_ms_per_sweep']))"
    done
  done
done
