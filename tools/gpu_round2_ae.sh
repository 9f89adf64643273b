#!/bin/bash
# round 2: roofline timed on the first-evaluation launches; default bench end to end
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2ag
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench default rc=$? lines=$(wc -l < $OUT/bench_default.json)"
timeout 600 python bench.py --steps 300 --warmup 5 --no-cpu --config cfg3 > $OUT/bench_cfg3.json 2> $OUT/bench_cfg3.err; echo "bench cfg3 rc=$?"
python - <<'PY'
import json
for f in ("default", "cfg3"):
    d = json.loads(open("gpurun_out/r2ag/bench_%s.json" % f).read())
    print(f, round(d["ms_per_step"], 4), round(d["value"] / 1e9, 3), d["roofline"], d["kernel_ms"]["per_evaluation"])
PY
