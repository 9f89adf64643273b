#!/bin/bash
# round 2: the tile's Gram matrix on the fp64 tensor cores (DMMA) in the fit kernel
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2aj
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
ST=$PWD/limo-velo_b200/liblimovelo_b200_st.so
LV_LIB_PATH=$ST timeout 200 python tools/step_timing.py > $OUT/step.log 2>&1; echo "step rc=$?"
LV_LIB_PATH=$ST timeout 200 python tools/timeline.py > $OUT/timeline.log 2>&1; echo "timeline rc=$?"
for c in cfg1 cfg3; do timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu --config $c > $OUT/bench_$c.json 2> $OUT/bench_$c.err; echo "bench $c rc=$?"; done
tail -n 3 $OUT/pytest.log
grep -o "FIT block 1.*group sum [0-9-]*" $OUT/step.log | tail -3
grep -A5 "update 4 (warm)" $OUT/timeline.log | tail -4 | sed 's/.*fit/fit/' | cut -c1-60
