#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2ac
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python tools/debug/cfg2_diff.py > $OUT/cfg2_diff.log 2>&1; echo "rc=$?"
grep -v "^Multi\|^Rebuild" $OUT/cfg2_diff.log | tail -30
