#!/bin/bash
# gpurun -- bash scripts/gpu_variants.sh <workload> lib[:bench args] ...   -- kernel-variant matrix: one `bench.py --quick` line per prebuilt
# library (make -C mcmc_colorer_b200/csrc OUT=$PWD/variants/libX.so EXTRA=-DMCMCB200_...=...), chosen through MCMCB200_LIB.
# Example: bash scripts/gpu_variants.sh c3 mcmc_colorer_b200/libmcmcb200.so variants/libKU10.so "variants/libKU10.so:--stage-cap-bytes 36864"
W=${1:-c3}; shift
mkdir -p gpurun_out; rm -f gpurun_out/variants.jsonl
for spec in "$@"; do
  lib=${spec%%:*}; extra=""; [ "$spec" != "$lib" ] && extra=${spec#*:}
  MCMCB200_LIB=$PWD/$lib timeout 300 python bench.py --workload $W --quick --steps 5 --warmup 3 $extra >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err
done
python - <<'PY'
import json
for l in open('gpurun_out/variants.jsonl'):
    d=json.loads(l); print(d['lib'].split('/')[-1], d['workload'], d['tuning'], d['kernel_mode'], round(d['ms_per_step'],3), round(d['chain_ms_per_sweep'],3), round(d['create_ms']), round(d['frac'],3))
PY
