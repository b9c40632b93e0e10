#!/bin/bash
# gpurun -- bash scripts/gpu_variants.sh lib:capBytes[:workload] ...   (prebuilt kernel variants under variants/, chosen via MCMCB200_LIB)
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read lib cap wl <<< "$spec"; wl=${wl:-c3}
  MCMCB200_LIB=$PWD/variants/lib_$lib.so MCMCB200_STAGE_CAP_BYTES=$cap timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/var_$lib_$cap.err | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.readline()); print('[$spec] ms %.3f frac %.3f chain %.3f launches/sweep %d' % (d['ms_per_step'], d['roofline']['frac'], d['chain_ms_per_sweep'], d['roofline']['launches_per_sweep']))
except Exception as e: print('[$spec] failed', e)"
done
