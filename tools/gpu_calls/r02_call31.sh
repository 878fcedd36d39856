#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python bench.py --steps 3 --warmup 3 2>gpurun_out/c31.err | tail -1 > gpurun_out/c31_default.json
python - <<PY
import json
r=json.loads(open("gpurun_out/c31_default.json").read())
def show(n, r):
  print(n, round(r["value"],1), round(r["ms_per_step"],2), r["clocks"]["sm_mhz"], r["roofline"]["frac"] if r.get("roofline") else None, "e2e", round(r["e2e"]["value"],1), r["gpu_launches"])
show("c4", r)
for k, v in r.get("extra", {}).items():
  if isinstance(v, dict) and "value" in v: show(k, v)
PY
for gb in 64; do
timeout 600 python bench.py --workload c4 --no-extras --no-cpu-baseline --global-batch $gb --steps 5 --warmup 3 2>>gpurun_out/c31.err | tail -1 > gpurun_out/c31_c4_$gb.json
python - <<PY
import json
r=json.loads(open("gpurun_out/c31_c4_$gb.json").read())
print("c4 gb=$gb", r["value"], r["ms_per_step"], r["clocks"]["sm_mhz"], r["roofline"]["frac"], "e2e", r["e2e"]["value"], r["gpu_launches"], r["config"]["execution"][:40])
PY
done
for gb in 32; do
timeout 600 python bench.py --workload c3 --no-extras --no-cpu-baseline --global-batch $gb --steps 8 --warmup 3 2>>gpurun_out/c31.err | tail -1 > gpurun_out/c31_c3_$gb.json
python - <<PY
import json
r=json.loads(open("gpurun_out/c31_c3_$gb.json").read())
print("c3 gb=$gb", r["value"], r["ms_per_step"], r["clocks"]["sm_mhz"], "e2e", r["e2e"]["value"], r["config"]["l2"][:60])
PY
done
MVB_GRAPH_MAX_ROWS=20000 timeout 600 python bench.py --workload c4 --no-extras --no-cpu-baseline --steps 3 --warmup 3 2>>gpurun_out/c31.err | tail -1 > gpurun_out/c31_c4_graph512.json
python - <<PY
import json
r=json.loads(open("gpurun_out/c31_c4_graph512.json").read())
print("c4 gb=512 forced graph", r["value"], r["ms_per_step"], r["clocks"]["sm_mhz"], r["roofline"]["frac"], "e2e", r["e2e"]["value"])
PY
tail -5 gpurun_out/c31.err
