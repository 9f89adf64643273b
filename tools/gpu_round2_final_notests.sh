SKIP_TESTS=1 bash tools/gpu_round2_final.sh
