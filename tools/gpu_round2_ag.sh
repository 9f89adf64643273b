#!/bin/bash
# round 2: phase clocks of one block of the fit kernel (tuning build)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2ai
mkdir -p $OUT
export PYTHONUNBUFFERED=1
ST=$PWD/limo-velo_b200/liblimovelo_b200_st.so
LV_LIB_PATH=$ST timeout 200 python tools/step_timing.py > $OUT/step.log 2>&1; echo "rc=$?"
grep -o "FIT block 1.*group sum [0-9-]*" $OUT/step.log | tail -6
