#!/bin/bash
# round 2, final measurement pass: tests, smoke, ncu launch list + --set full captures, the benches of every config
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2final
mkdir -p $OUT
export PYTHONUNBUFFERED=1
: > $OUT/summary.txt
say() { echo "$@" | tee -a $OUT/summary.txt; }
if [ "${SKIP_TESTS:-0}" != 1 ]; then timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; say "pytest rc=$?"; fi
timeout 300 python -c "import __graft_entry__ as G; G.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; say "smoke rc=$?"
# launch list (one line per launch, serialised, cold caches): shares of the step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > $OUT/launches_bench.log 2>&1; say "ncu launches rc=$?"
# full captures (~2 MB per launch; gpurun brings back at most 64 MiB): eval 0, eval 1 and part of eval 2 of one update; the kernels of one map update
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^lv_(search|fit|ieskf_step|reuse)' -s 38 -c 12 \
    -o $OUT/prof_update python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/ncu_update.log 2>&1; say "ncu update rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lv_map_|lv_sweep_to_world|DeviceRadixSort' -s 30 -c 10 \
    -o $OUT/prof_map python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/ncu_map.log 2>&1; say "ncu map rc=$?"
ST=$PWD/limo-velo_b200/liblimovelo_b200_st.so
LV_LIB_PATH=$ST timeout 200 python tools/timeline.py > $OUT/timeline.log 2>&1; say "timeline rc=$?"
b() { local name=$1; shift; timeout 900 python bench.py "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err; say "bench $name rc=$?"; }
b cfg1 --sequences-per-gpu 1,2,4,8
b cfg1_reference --impl reference --steps 20 --warmup 3
b cfg0 --config cfg0 --steps 300 --warmup 5 --no-cpu
b cfg2 --config cfg2 --steps 300 --warmup 5 --no-cpu
b cfg3 --config cfg3 --steps 300 --warmup 5 --no-cpu
b cfg4 --config cfg4 --steps 300 --warmup 5 --no-cpu
b cfg1_sorted --config cfg1 --sort-queries 1 --steps 300 --warmup 5 --no-cpu
b cfg1_v06 --config cfg1 --voxel 0.6 --steps 300 --warmup 5 --no-cpu
du -sh $OUT
tail -n 3 $OUT/bench_*.err
