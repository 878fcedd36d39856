#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -5
echo "== graph auto"; timeout 300 python tools/gpu_probe_e2e.py 64 2>&1 | tail -8
for gb in 64 128; do
timeout 600 python bench.py --workload c4 --no-extras --global-batch $gb --steps 5 --warmup 3 2>gpurun_out/c27.err | tail -1 > gpurun_out/c27_c4_$gb.json
python - <<PY
import json
r=json.loads(open("gpurun_out/c27_c4_$gb.json").read())
print("c4 gb=$gb", r["value"], r["ms_per_step"], r["clocks"]["sm_mhz"], r["roofline"]["frac"], "e2e", r["e2e"]["value"])
PY
done
