#!/bin/bash
# round 2, call A: (1) round-1 search kernel under other code generations, (2) GPU test suite on the incremental map,
# (3) benches of all configs
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2a
mkdir -p $OUT
export PYTHONUNBUFFERED=1
: > $OUT/summary.txt
for v in r1diag r1o1 r1noldg; do
  timeout 100 python tools/k1_isolate_raw.py tools/r1_libs/liblimovelo_b200_$v.so 0.35 50 5 7 0 > $OUT/iso_$v.log 2>&1
  echo "iso $v rc=$?" | tee -a $OUT/summary.txt
done
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt
for c in cfg1 cfg0 cfg2 cfg3; do
  timeout 420 python bench.py --config $c --steps 200 --warmup 5 --no-cpu > $OUT/bench_$c.json 2> $OUT/bench_$c.err
  echo "bench $c rc=$?" | tee -a $OUT/summary.txt
done
tail -n 4 $OUT/iso_*.log
tail -n 40 $OUT/pytest_gpu.log
tail -n 3 $OUT/bench_*.err
