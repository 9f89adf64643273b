#!/bin/bash
# round 2, call AC: is the ring search paying for a cold instruction cache?  The same kernel once, twice, four times in a row (kernel_ms of the bench's profile pass)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2ae
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for n in 0 1 3; do
  for c in cfg1 cfg2; do
    LV_RINGS_AGAIN=$n timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --config $c > $OUT/bench_${c}_again$n.json 2> $OUT/bench_${c}_again$n.err; echo "again=$n $c rc=$?"
  done
done
