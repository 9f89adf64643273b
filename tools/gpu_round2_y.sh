#!/bin/bash
# round 2, call Y: wider margin of the ring search (reusable answers): hard counts per evaluation, benches
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2ab
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_voxel_sweep.py tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
ST=$PWD/limo-velo_b200/liblimovelo_b200_st.so
for c in cfg1 cfg2 cfg3; do
  LV_TIMELINE_CFG=$c LV_LIB_PATH=$ST timeout 300 python tools/timeline.py > $OUT/timeline_$c.log 2>&1; echo "timeline $c rc=$?"
  grep -A5 "update 4 (warm)" $OUT/timeline_$c.log | tail -4 | sed 's/.*upper/upper/' | cut -c1-60,150-260
  timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu --config $c > $OUT/bench_$c.json 2> $OUT/bench_$c.err; echo "bench $c rc=$?"
done
tail -n 2 $OUT/pytest.log
