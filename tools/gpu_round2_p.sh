#!/bin/bash
# round 2, call P: 4 lanes per query in the level-0 search (the instantiation ptxas miscompiled on the round-1 sources), every step under a timeout
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2s
mkdir -p $OUT
export PYTHONUNBUFFERED=1
: > $OUT/summary.txt
say() { echo "$@" | tee -a $OUT/summary.txt; }
export LV_SEARCH_GROUP=4
timeout 400 python -m pytest tests/test_gpu_voxel_sweep.py -m gpu -q -p no:cacheprovider -x > $OUT/pytest_sweep.log 2>&1; say "pytest voxel sweep g4 rc=$?"
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x --deselect tests/test_gpu_voxel_sweep.py > $OUT/pytest_rest.log 2>&1; say "pytest rest g4 rc=$?"
ST=$PWD/limo-velo_b200/liblimovelo_b200_st.so
LV_LIB_PATH=$ST timeout 200 python tools/timeline.py > $OUT/timeline_g4.log 2>&1; say "timeline g4 rc=$?"
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu > $OUT/bench_cfg1_g4.json 2> $OUT/bench_cfg1_g4.err; say "bench cfg1 g4 rc=$?"
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu --config cfg3 > $OUT/bench_cfg3_g4.json 2> $OUT/bench_cfg3_g4.err; say "bench cfg3 g4 rc=$?"
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu --config cfg2 > $OUT/bench_cfg2_g4.json 2> $OUT/bench_cfg2_g4.err; say "bench cfg2 g4 rc=$?"
tail -n 3 $OUT/pytest_sweep.log $OUT/pytest_rest.log
grep -A5 "update 4 (warm)" $OUT/timeline_g4.log | tail -5
nvidia-smi --query-gpu=name,utilization.gpu --format=csv
