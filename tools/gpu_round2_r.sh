#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2u
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python tools/debug/stream_last.py > $OUT/stream_last.log 2>&1; echo "rc=$?"
LV_SEARCH_GROUP=8 timeout 600 python tools/debug/stream_last.py > $OUT/stream_last_g8.log 2>&1; echo "g8 rc=$?"
grep -v "^Multi\|^Rebuild" $OUT/stream_last.log | tail -60
echo ==== g8
grep -v "^Multi\|^Rebuild" $OUT/stream_last_g8.log | grep "^sweep"
