#!/bin/bash
# round 2, call G: ncu on the ring-search kernel of a first evaluation, then compute-sanitizer over the whole path
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2j
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'lv_search_rings_kernel' -s 0 -c 6 \
    -o $OUT/prof_rings python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_rings.log 2>&1
echo "ncu rings rc=$?" | tee $OUT/summary.txt
bash tools/gpu_sanitize.sh
cat gpurun_out/sanitize/summary.txt | tee -a $OUT/summary.txt
