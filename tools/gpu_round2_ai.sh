#!/bin/bash
# round 2: racecheck + memcheck + synccheck on the fit kernel's tensor-core fold (quick target: two search shapes)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/sanitize2
mkdir -p $OUT
export PYTHONUNBUFFERED=1 LV_SANITIZE_QUICK=1
for tool in racecheck memcheck synccheck initcheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > $OUT/$tool.log 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/$tool.log | tail -1)"
done
