#!/bin/bash
# round 2, call L: tests + timeline + benches after the two-stage blind search
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2o
mkdir -p $OUT
export PYTHONUNBUFFERED=1
: > $OUT/summary.txt
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt
ST=$PWD/limo-velo_b200/liblimovelo_b200_st.so
LV_LIB_PATH=$ST timeout 200 python tools/timeline.py > $OUT/timeline_s0.log 2>&1; echo "timeline s0 rc=$?" | tee -a $OUT/summary.txt
b() { local name=$1; shift; timeout 600 python bench.py --steps 300 --warmup 5 --no-cpu "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo "bench $name rc=$?" | tee -a $OUT/summary.txt; }
b cfg1_v04 --config cfg1
b cfg1_v06 --config cfg1 --voxel 0.6
b cfg2 --config cfg2
b cfg3 --config cfg3
tail -n 5 $OUT/pytest_gpu.log
tail -n 12 $OUT/timeline_s0.log
tail -n 3 $OUT/bench_*.err
