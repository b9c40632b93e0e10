#!/bin/bash
# round 2, call A: full GPU test suite + smoke + first timings of the TMA-staged blocked sweep on config 3
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2a_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2a_smoke.log
tail -2 gpurun_out/r2a_smoke.log
for v in "" "--no-overlap" "--stage-cap-bytes 22528 --stage-buffers 2" "--stage-cap-bytes 32768" "--item-bits 17" "--item-bits 19"; do
  timeout 300 python bench.py --workload c3 --quick --steps 5 --warmup 3 $v >> gpurun_out/r2a_quick.jsonl 2>> gpurun_out/r2a_quick.err
done
cat gpurun_out/r2a_quick.jsonl
