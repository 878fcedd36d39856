#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
for spec in "c3 32" "c3 256" "c4 64"; do
set -- $spec
timeout 600 python bench.py --workload $1 --no-extras --no-cpu-baseline --global-batch $2 --steps 8 --warmup 3 2>>gpurun_out/c40.err | tail -1 > gpurun_out/c40_$1_$2.json
python - <<PY
import json
r=json.loads(open("gpurun_out/c40_$1_$2.json").read())
print("$1 gb=$2", round(r["value"],1), round(r["ms_per_step"],2), r["clocks"]["sm_mhz"], "e2e", round(r["e2e"]["value"],1))
PY
done
