#!/bin/bash
# round 2, call Q: level-0 search shape chosen per evaluation (4 lanes per query for the big launches, 8 for the short lists)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2t
mkdir -p $OUT
export PYTHONUNBUFFERED=1
: > $OUT/summary.txt
say() { echo "$@" | tee -a $OUT/summary.txt; }
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $OUT/pytest.log 2>&1; say "pytest rc=$?"
ST=$PWD/limo-velo_b200/liblimovelo_b200_st.so
LV_LIB_PATH=$ST timeout 200 python tools/timeline.py > $OUT/timeline.log 2>&1; say "timeline rc=$?"
for c in cfg1 cfg0 cfg2 cfg3; do
  timeout 400 python bench.py --steps 300 --warmup 5 --no-cpu --config $c > $OUT/bench_$c.json 2> $OUT/bench_$c.err; say "bench $c rc=$?"
done
LV_SEARCH_GROUP=8 timeout 400 python bench.py --steps 300 --warmup 5 --no-cpu > $OUT/bench_cfg1_g8.json 2> $OUT/bench_cfg1_g8.err; say "bench cfg1 g8 rc=$?"
tail -n 3 $OUT/pytest.log
grep -A5 "update 4 (warm)" $OUT/timeline.log | tail -5 | cut -c1-125
grep -A5 "update 5 (flushed" $OUT/timeline.log | tail -5 | cut -c1-125
