#!/bin/bash
# round 2, call C: tests, then cfg1 at both voxel sizes with the reworked ring search, several sequences per GPU, the other configs
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2c
mkdir -p $OUT
export PYTHONUNBUFFERED=1
: > $OUT/summary.txt
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt
b() { local name=$1; shift; timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo "bench $name rc=$?" | tee -a $OUT/summary.txt; }
b cfg1_v04 --config cfg1 --sequences-per-gpu 1,2,4,8
b cfg1_v06 --config cfg1 --voxel 0.6
b cfg1_s1 --config cfg1 --sort-queries 1
b cfg2 --config cfg2
b cfg3 --config cfg3
b cfg0 --config cfg0
tail -n 30 $OUT/pytest_gpu.log
tail -n 3 $OUT/bench_*.err
