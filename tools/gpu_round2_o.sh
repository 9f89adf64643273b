#!/bin/bash
# round 2, call O: the query-per-lane search kernel (LV_SEARCH_GROUP=32) against the default: tests, timelines, benches
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2r
mkdir -p $OUT
export PYTHONUNBUFFERED=1
: > $OUT/summary.txt
say() { echo "$@" | tee -a $OUT/summary.txt; }
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_voxel_sweep.py -m gpu -q -p no:cacheprovider -x > $OUT/pytest_a.log 2>&1; say "pytest default rc=$?"
LV_SEARCH_GROUP=32 timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $OUT/pytest_g32.log 2>&1; say "pytest group32 rc=$?"
ST=$PWD/limo-velo_b200/liblimovelo_b200_st.so
for g in 8 32; do
  LV_SEARCH_GROUP=$g LV_LIB_PATH=$ST timeout 200 python tools/timeline.py > $OUT/timeline_g$g.log 2>&1; say "timeline g$g rc=$?"
  LV_SEARCH_GROUP=$g timeout 600 python bench.py --steps 300 --warmup 5 --no-cpu > $OUT/bench_cfg1_g$g.json 2> $OUT/bench_cfg1_g$g.err; say "bench cfg1 g$g rc=$?"
  LV_SEARCH_GROUP=$g timeout 600 python bench.py --steps 300 --warmup 5 --no-cpu --config cfg3 > $OUT/bench_cfg3_g$g.json 2> $OUT/bench_cfg3_g$g.err; say "bench cfg3 g$g rc=$?"
done
tail -n 3 $OUT/pytest_a.log $OUT/pytest_g32.log
for g in 8 32; do echo "== g$g"; grep -A5 "update 4 (warm)" $OUT/timeline_g$g.log | tail -5; done
