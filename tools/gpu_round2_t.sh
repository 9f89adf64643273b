#!/bin/bash
# round 2, call T: eight loads in flight per lane in the 4-lane search
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2w
mkdir -p $OUT
export PYTHONUNBUFFERED=1
: > $OUT/summary.txt
say() { echo "$@" | tee -a $OUT/summary.txt; }
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_voxel_sweep.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x > $OUT/pytest.log 2>&1; say "pytest rc=$?"
ST=$PWD/limo-velo_b200/liblimovelo_b200_st.so
LV_LIB_PATH=$ST timeout 200 python tools/timeline.py > $OUT/timeline.log 2>&1; say "timeline rc=$?"
for c in cfg1 cfg3; do
  timeout 400 python bench.py --steps 300 --warmup 5 --no-cpu --config $c > $OUT/bench_$c.json 2> $OUT/bench_$c.err; say "bench $c rc=$?"
done
timeout 400 python bench.py --steps 300 --warmup 5 --no-cpu > $OUT/bench_cfg1_b.json 2> $OUT/bench_cfg1_b.err; say "bench cfg1 again rc=$?"
tail -n 3 $OUT/pytest.log
grep -A5 "update 4 (warm)" $OUT/timeline.log | tail -5 | cut -c1-125
grep -A5 "update 5 (flushed" $OUT/timeline.log | tail -5 | cut -c1-125
