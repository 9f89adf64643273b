#!/bin/bash
# round 2, call K: ncu on the ring-search kernel (first evaluation) after the rank-dealing rewrite
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2n
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'lv_search_rings_kernel' -s 4 -c 1 \
    -o $OUT/prof_rings2 python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_rings.log 2>&1
echo "ncu rings rc=$?" | tee $OUT/summary.txt
