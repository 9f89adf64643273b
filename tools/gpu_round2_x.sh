#!/bin/bash
# round 2, call X: where the time of the small configs goes (timelines of cfg2, cfg0, cfg3)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2aa
mkdir -p $OUT
export PYTHONUNBUFFERED=1
ST=$PWD/limo-velo_b200/liblimovelo_b200_st.so
for c in cfg1 cfg2 cfg3; do
  LV_TIMELINE_CFG=$c LV_LIB_PATH=$ST timeout 300 python tools/timeline.py > $OUT/timeline_$c.log 2>&1; echo "timeline $c rc=$?"
  grep -A5 "update 4 (warm)" $OUT/timeline_$c.log | tail -5 | cut -c150-260
done
