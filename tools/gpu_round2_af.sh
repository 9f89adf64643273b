#!/bin/bash
# round 2: the C++ driver (reference's tick on the device) once more on the final library
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2ah
mkdir -p $OUT
timeout 300 ./limo-velo_b200/limovelo_synth limo-velo_b200/config/xaloc.yaml 8 200000 64 1024 > $OUT/synth.log 2>&1; echo "synth rc=$?"
timeout 300 ./limo-velo_b200/limovelo_synth limo-velo_b200/config/xaloc.yaml 4 200000 64 1024 0.3 > $OUT/synth_leaf.log 2>&1; echo "synth leaf rc=$?"
tail -n 9 $OUT/synth.log; tail -n 5 $OUT/synth_leaf.log
