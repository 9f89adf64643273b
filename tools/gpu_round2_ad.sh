#!/bin/bash
# round 2: the bench's stdout under torchrun and alone (exactly one line, JSON)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2af
mkdir -p $OUT
export PYTHONUNBUFFERED=1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29521 bench.py --gpus 2 --steps 100 --warmup 3 --no-cpu > $OUT/n2.out 2> $OUT/n2.err; echo "n2 rc=$? lines=$(wc -l < $OUT/n2.out)"
timeout 600 $T --master-port 29522 bench.py --gpus 2 --steps 2 --warmup 3 --impl reference > $OUT/n2ref.out 2> $OUT/n2ref.err; echo "n2 ref rc=$? lines=$(wc -l < $OUT/n2ref.out)"
timeout 600 python bench.py --steps 100 --warmup 3 --no-cpu > $OUT/n1.out 2> $OUT/n1.err; echo "n1 rc=$? lines=$(wc -l < $OUT/n1.out)"
timeout 600 $T --master-port 29523 bench.py --gpus 2 --steps 100 --warmup 3 --no-cpu --config cfg4 > $OUT/n2cfg4.out 2> $OUT/n2cfg4.err; echo "n2 cfg4 rc=$? lines=$(wc -l < $OUT/n2cfg4.out)"
python - <<'PY'
import json
for f in ("n2", "n2ref", "n1", "n2cfg4"):
    d = json.loads(open("gpurun_out/r2af/%s.out" % f).read())
    print(f, d.get("impl", "native"), d["n_gpus"], round(d["ms_per_step"], 4), round(d["value"] / 1e9, 4))
PY
grep -c "NCCL version" $OUT/n2.err
