#!/bin/bash
# round 2: the per-evaluation choice of the search shape against 4 lanes everywhere (after the 8-loads-in-flight change), two runs each
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2ak
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for r in 1 2; do
  timeout 300 python bench.py --steps 400 --warmup 5 --no-cpu > $OUT/bench_default_$r.json 2> /dev/null; echo "default $r rc=$?"
  LV_SEARCH_GROUP=4 timeout 300 python bench.py --steps 400 --warmup 5 --no-cpu > $OUT/bench_g4_$r.json 2> /dev/null; echo "g4 $r rc=$?"
done
python - <<'PY'
import json
for f in ("default_1", "g4_1", "default_2", "g4_2"):
    d = json.loads(open("gpurun_out/r2ak/bench_%s.json" % f).read())
    print(f, round(d["ms_per_step"], 4), round(d["e2e"]["ms_per_step"], 4))
PY
