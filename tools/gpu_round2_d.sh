#!/bin/bash
# round 2, call D: graph-mode timeline of one update (both search paths), ncu launch lists and --set full captures
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2d
mkdir -p $OUT
export PYTHONUNBUFFERED=1
: > $OUT/summary.txt
ST=$PWD/limo-velo_b200/liblimovelo_b200_st.so
LV_LIB_PATH=$ST timeout 200 python tools/timeline.py > $OUT/timeline_s0.log 2>&1; echo "timeline s0 rc=$?" | tee -a $OUT/summary.txt
LV_LIB_PATH=$ST LV_TIMELINE_SORT=1 timeout 200 python tools/timeline.py > $OUT/timeline_s1.log 2>&1; echo "timeline s1 rc=$?" | tee -a $OUT/summary.txt
LV_LIB_PATH=$ST timeout 200 python tools/step_timing.py > $OUT/step_timing.log 2>&1; echo "step_timing rc=$?" | tee -a $OUT/summary.txt
for s in 0 1; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_s$s.csv \
      python bench.py --steps 4 --warmup 3 --no-cpu --sort-queries $s > $OUT/ncu_launch_s$s.log 2>&1
  echo "ncu launches s$s rc=$?" | tee -a $OUT/summary.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lv_search_kernel|lv_search_rings_kernel|lv_fit_kernel|lv_ieskf_step_kernel' -s 8 -c 8 \
    -o $OUT/prof_s0 python bench.py --steps 3 --warmup 3 --no-cpu --sort-queries 0 > $OUT/ncu_full_s0.log 2>&1
echo "ncu full s0 rc=$?" | tee -a $OUT/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lv_search_staged_kernel|lv_bin_kernel' -s 4 -c 5 \
    -o $OUT/prof_s1 python bench.py --steps 3 --warmup 3 --no-cpu --sort-queries 1 > $OUT/ncu_full_s1.log 2>&1
echo "ncu full s1 rc=$?" | tee -a $OUT/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lv_map_' -s 8 -c 8 \
    -o $OUT/prof_map python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/ncu_full_map.log 2>&1
echo "ncu full map rc=$?" | tee -a $OUT/summary.txt
ls -la $OUT
tail -n 30 $OUT/timeline_s0.log
