#!/bin/bash
# round 2, call V: ring search with its serial fallback out of line
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2y
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_voxel_sweep.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
ST=$PWD/limo-velo_b200/liblimovelo_b200_st.so
LV_LIB_PATH=$ST timeout 200 python tools/timeline.py > $OUT/timeline.log 2>&1; echo "timeline rc=$?"
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu > $OUT/bench_cfg1.json 2> $OUT/bench_cfg1.err; echo "bench rc=$?"
tail -n 2 $OUT/pytest.log
grep -A2 "update 4 (warm)" $OUT/timeline.log | tail -1 | cut -c1-140
grep -A2 "update 5 (flushed" $OUT/timeline.log | tail -1 | cut -c1-140
