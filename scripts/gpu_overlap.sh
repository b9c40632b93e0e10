#!/bin/bash
# gpu tests, then config 3 with pass A || pass B overlapped (default) and serial (MCMCB200_NO_OVERLAP=1), for each stage cap given
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -rs --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
one() {
  timeout 300 python bench.py --workload ${WL:-c3} --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/ov.err | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.readline()); print('[$1] ms %.3f frac %.3f chain %.3f' % (d['ms_per_step'], d['roofline']['frac'], d['chain_ms_per_sweep']))
except Exception as e: print('[$1] failed', e)"; tail -2 gpurun_out/ov.err
}
for cap in "$@"; do
  MCMCB200_STAGE_CAP_BYTES=$cap one "overlap cap=$cap"
  MCMCB200_STAGE_CAP_BYTES=$cap MCMCB200_NO_OVERLAP=1 one "serial  cap=$cap"
done
