#!/bin/bash
mkdir -p gpurun_out
one() {
  timeout 300 python bench.py --workload ${WL:-c3} --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/ov.err | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.readline()); print('[$1] ms %.3f frac %.3f chain %.3f' % (d['ms_per_step'], d['roofline']['frac'], d['chain_ms_per_sweep']))
except Exception as e: print('[$1] failed', e)"; tail -2 gpurun_out/ov.err
}
export MCMCB200_STAGE_CAP_BYTES=45056
one "overlap A1 B2"
MCMCB200_B_PER_SM=1 one "overlap B1 (A up to 3)"
MCMCB200_B_PER_SM=1 MCMCB200_A_PER_SM=2 one "overlap B1 A2"
MCMCB200_STAGE_CAP_BYTES=65504 MCMCB200_B_PER_SM=1 one "overlap B1 cap64K"
MCMCB200_NO_OVERLAP=1 MCMCB200_A_PER_SM=2 one "serial A2 B2"
