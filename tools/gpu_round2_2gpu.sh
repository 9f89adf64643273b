#!/bin/bash
# round 2: two GPUs of one box — weak scaling of cfg1 (one sequence per GPU) and the fixed 8-sequence job cfg4
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2gpu2
mkdir -p $OUT
export PYTHONUNBUFFERED=1
: > $OUT/summary.txt
say() { echo "$@" | tee -a $OUT/summary.txt; }
nvidia-smi -L > $OUT/gpus.txt 2>&1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 5 --no-cpu > $OUT/bench_cfg1_n2.json 2> $OUT/bench_cfg1_n2.err; say "cfg1 n2 rc=$?"
timeout 600 $T --master-port 29512 bench.py --gpus 2 --steps 300 --warmup 5 --no-cpu --config cfg4 > $OUT/bench_cfg4_n2.json 2> $OUT/bench_cfg4_n2.err; say "cfg4 n2 rc=$?"
timeout 600 $T --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --impl reference > $OUT/bench_ref_n2.json 2> $OUT/bench_ref_n2.err; say "reference n2 rc=$?"
timeout 600 python bench.py --gpus 1 --steps 300 --warmup 5 --no-cpu > $OUT/bench_cfg1_n1.json 2> $OUT/bench_cfg1_n1.err; say "cfg1 n1 rc=$?"
timeout 600 python bench.py --gpus 1 --steps 300 --warmup 5 --no-cpu --config cfg4 > $OUT/bench_cfg4_n1.json 2> $OUT/bench_cfg4_n1.err; say "cfg4 n1 rc=$?"
tail -n 3 $OUT/*.err
