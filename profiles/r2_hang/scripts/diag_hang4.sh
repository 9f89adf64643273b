#!/bin/bash
# fourth pass: isolate the hanging kernel in-process (tools/k1_isolate.py), then baseline benches of all configs
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/diag4
mkdir -p $OUT
export PYTHONUNBUFFERED=1
: > $OUT/summary.txt
D=$PWD/limo-velo_b200/liblimovelo_b200_diag.so
iso() { local name=$1; shift; env LV_LIB_PATH=$D "$@" > $OUT/$name.log 2>&1; echo "$name rc=$?" | tee -a $OUT/summary.txt; }
iso k1_g4      timeout 120 python tools/k1_isolate.py 0.35 1 100
iso k1_g4_b    timeout 120 python tools/k1_isolate.py 0.35 1 100
iso k1_g1      env LV_SEARCH_GROUP=1 timeout 120 python tools/k1_isolate.py 0.35 1 100
iso k1_g8      env LV_SEARCH_GROUP=8 timeout 120 python tools/k1_isolate.py 0.35 1 100
iso k12_g4     timeout 120 python tools/k1_isolate.py 0.35 2 100
iso k123_g4    timeout 120 python tools/k1_isolate.py 0.35 3 100
iso k123_g1    env LV_SEARCH_GROUP=1 timeout 120 python tools/k1_isolate.py 0.35 3 100
iso k123_v05   timeout 120 python tools/k1_isolate.py 0.5 3 100
for c in cfg0 cfg2 cfg3 cfg1; do
  timeout 420 python bench.py --config $c --steps 200 --warmup 5 --no-cpu > $OUT/bench_$c.json 2> $OUT/bench_$c.err
  echo "bench $c rc=$?" | tee -a $OUT/summary.txt
done
tail -n 12 $OUT/k1*.log
head -c 1500 $OUT/bench_cfg*.json
tail -n 5 $OUT/bench_*.err
