#!/bin/bash
# which blocked mode is faster for partition-sized graphs (n = 12.5M and 25M of the config-3 family)?
mkdir -p gpurun_out
one() {
  timeout 300 python bench.py --workload c3 --n $N --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/ps.err | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.readline()); print('[$1 n=$N] ms %.3f frac %.3f chain %.3f mode %s' % (d['ms_per_step'], d['roofline']['frac'], d['chain_ms_per_sweep'], d['roofline']['kernel_mode']))
except Exception as e: print('[$1] failed', e)"; tail -2 gpurun_out/ps.err
}
for N in 12500000 25000000 50000000; do
  one "default"
  MCMCB200_STAGE_CAP_BYTES=45056 MCMCB200_ITEM_BITS=18 one "overlap i18"
  MCMCB200_STAGE_CAP_BYTES=45056 MCMCB200_ITEM_BITS=17 one "overlap i17"
done
