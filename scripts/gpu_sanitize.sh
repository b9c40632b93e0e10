#!/bin/bash
# compute-sanitizer over every kernel family: memcheck, racecheck (shared-memory hazards: walk queues, stage
# buffers, bitmaps), synccheck.  Logs -> gpurun_out/sanitize_*.log (copied to profiles/ by hand).
mkdir -p gpurun_out
python scripts/sanitize_driver.py > gpurun_out/sanitize_plain.log 2>&1 || { tail -5 gpurun_out/sanitize_plain.log; exit 1; }
for tool in memcheck racecheck synccheck; do
  SANITIZE_N=${SANITIZE_N:-6000} timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_driver.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=|all ok" gpurun_out/sanitize_$tool.log
done
