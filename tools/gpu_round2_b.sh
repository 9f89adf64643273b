#!/bin/bash
# round 2, call B: GPU test suite, then cfg1 with both search paths and both voxel sizes, then the other configs
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2b
mkdir -p $OUT
export PYTHONUNBUFFERED=1
: > $OUT/summary.txt
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt
b() { local name=$1; shift; timeout 420 python bench.py --steps 200 --warmup 5 --no-cpu "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo "bench $name rc=$?" | tee -a $OUT/summary.txt; }
b cfg1_s0 --config cfg1 --sort-queries 0
b cfg1_s1 --config cfg1 --sort-queries 1
b cfg1_s0_v06 --config cfg1 --sort-queries 0 --voxel 0.6
b cfg1_s1_v06 --config cfg1 --sort-queries 1 --voxel 0.6
b cfg2_s0 --config cfg2 --sort-queries 0
b cfg2_s1 --config cfg2 --sort-queries 1
b cfg3_s0 --config cfg3 --sort-queries 0
b cfg3_s1 --config cfg3 --sort-queries 1
tail -n 30 $OUT/pytest_gpu.log
tail -n 3 $OUT/bench_*.err
