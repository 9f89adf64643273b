#!/bin/bash
# round 2, call W: the pivot's reciprocal computed beside the row swap in the 12x25 elimination
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2z
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
ST=$PWD/limo-velo_b200/liblimovelo_b200_st.so
LV_LIB_PATH=$ST timeout 200 python tools/step_timing.py > $OUT/step.log 2>&1; echo "step_timing rc=$?"
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu > $OUT/bench_cfg1.json 2> $OUT/bench_cfg1.err; echo "bench rc=$?"
tail -n 2 $OUT/pytest.log
tail -n 4 $OUT/step.log
