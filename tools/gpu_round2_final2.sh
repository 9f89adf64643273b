#!/bin/bash
# round 2, last pass: sanitizers on the final code, then the final measurement script
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
bash tools/gpu_sanitize.sh
bash tools/gpu_round2_final.sh
cat gpurun_out/sanitize/summary.txt
