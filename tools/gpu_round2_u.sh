#!/bin/bash
# round 2, call U: 2 lanes per query, for the record
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2x
mkdir -p $OUT
export PYTHONUNBUFFERED=1
export LV_SEARCH_GROUP=2
timeout 300 python -m pytest tests/test_gpu_voxel_sweep.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider -x > $OUT/pytest.log 2>&1; echo "pytest g2 rc=$?"
ST=$PWD/limo-velo_b200/liblimovelo_b200_st.so
LV_LIB_PATH=$ST timeout 200 python tools/timeline.py > $OUT/timeline.log 2>&1; echo "timeline rc=$?"
for c in cfg1 cfg3; do timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu --config $c > $OUT/bench_$c.json 2> $OUT/bench_$c.err; echo "bench $c rc=$?"; done
tail -n 2 $OUT/pytest.log
grep -A5 "update 4 (warm)" $OUT/timeline.log | tail -5 | cut -c1-100
grep -A5 "update 5 (flushed" $OUT/timeline.log | tail -5 | cut -c1-100
