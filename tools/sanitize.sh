#!/bin/bash
# compute-sanitizer over every scan-kernel mode: memcheck, racecheck (shared-memory hazards: the counter rows,
# fold_sums, the per-CTA floor, the mbarrier-staged key buffers), synccheck.  Logs land in gpurun_out/<tag>_sanitizer_*.log.
set -u
tag=${1:-r02}; out=gpurun_out; mkdir -p $out
export KTA_NO_BUILD=1
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --num-cuda-barriers 32768 --print-limit 20 --error-exitcode 77 python tools/sanitize_driver.py \
      > $out/${tag}_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $out/${tag}_sanitizer_summary.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error" $out/${tag}_sanitizer_$tool.log | head -5 | tee -a $out/${tag}_sanitizer_summary.log
done
