#!/bin/bash
# quick kernel experiments: rebuild with EXTRA flags on the box, run the c5/c3 sweeps
mkdir -p gpurun_out
run() {
  make -s -C mcmc_colorer_b200/csrc clean; make -s -C mcmc_colorer_b200/csrc EXTRA="$1" > /dev/null 2>&1
  for W in c5 c3; do
    timeout 300 python bench.py --workload $W --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('[$1] $W ms %.3f frac %.3f chain %.3f' % (d['ms_per_step'], d['roofline']['frac'], d['chain_ms_per_sweep']))"
  done
}
for F in "$@"; do run "$F"; done
make -s -C mcmc_colorer_b200/csrc clean; make -s -C mcmc_colorer_b200/csrc > /dev/null 2>&1
