#!/bin/bash
# round 2, call N: smaller code in the ring-search and step kernels (instruction-cache misses): tests, timelines, step phases, bench
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2q
mkdir -p $OUT
export PYTHONUNBUFFERED=1
: > $OUT/summary.txt
say() { echo "$@" | tee -a $OUT/summary.txt; }
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $OUT/pytest_gpu.log 2>&1; say "pytest rc=$?"
for v in st stregs; do
  L=$PWD/limo-velo_b200/liblimovelo_b200_$v.so
  LV_LIB_PATH=$L timeout 200 python tools/timeline.py > $OUT/timeline_$v.log 2>&1; say "timeline $v rc=$?"
  LV_LIB_PATH=$L timeout 200 python tools/step_timing.py > $OUT/step_$v.log 2>&1; say "step_timing $v rc=$?"
done
timeout 600 python bench.py --steps 300 --warmup 5 --no-cpu > $OUT/bench_cfg1.json 2> $OUT/bench_cfg1.err; say "bench rc=$?"
tail -n 3 $OUT/pytest_gpu.log
for v in st stregs; do echo "== $v"; grep -A5 "update 4 (warm)" $OUT/timeline_$v.log | tail -5; tail -n 6 $OUT/step_$v.log; done
