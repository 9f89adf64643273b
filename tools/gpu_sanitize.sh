#!/bin/bash
# compute-sanitizer over every kernel of the path (tools/sanitize_target.py); logs -> gpurun_out/sanitize/
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/sanitize
mkdir -p $OUT
export PYTHONUNBUFFERED=1
: > $OUT/summary.txt
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > $OUT/$tool.log 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/$tool.log | tail -1)" | tee -a $OUT/summary.txt
done
