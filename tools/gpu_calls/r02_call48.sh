#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py 2>gpurun_out/c48.err | tail -1 > gpurun_out/c48_default.json
python - <<PY
import json
r=json.loads(open("gpurun_out/c48_default.json").read())
def show(n, r):
  print(n, round(r["value"],1), round(r["ms_per_step"],2), r["clocks"]["sm_mhz"], r["clocks"].get("power_w"), r["roofline"]["frac"] if r.get("roofline") else None, "e2e", round(r["e2e"]["value"],1), r["gpu_launches"])
show("c4", r)
for k, v in r.get("extra", {}).items():
  if isinstance(v, dict) and "value" in v: show(k, v)
PY
