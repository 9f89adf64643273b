#!/bin/bash
# round 2, call I: diagnosis of the evaluation-3 mismatch in the streaming test
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2l
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python tools/debug/stream_eval.py > $OUT/stream_eval.log 2>&1
echo "rc=$?"
grep -v "^Multi" $OUT/stream_eval.log | tail -80
