#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -5
for gb in 64 512; do
timeout 600 python bench.py --workload c4 --no-extras --no-cpu-baseline --global-batch $gb --steps 5 --warmup 3 2>gpurun_out/c30.err | tail -1 > gpurun_out/c30_c4_$gb.json
python - <<PY
import json
r=json.loads(open("gpurun_out/c30_c4_$gb.json").read())
print("c4 gb=$gb", r["value"], r["ms_per_step"], r["clocks"]["sm_mhz"], r["roofline"]["frac"], "e2e", r["e2e"]["value"])
PY
done
for gb in 32 256; do
timeout 600 python bench.py --workload c3 --no-extras --no-cpu-baseline --global-batch $gb --steps 8 --warmup 3 2>gpurun_out/c30.err | tail -1 > gpurun_out/c30_c3_$gb.json
python - <<PY
import json
r=json.loads(open("gpurun_out/c30_c3_$gb.json").read())
print("c3 gb=$gb", r["value"], r["ms_per_step"], r["clocks"]["sm_mhz"], "e2e", r["e2e"]["value"])
PY
done
