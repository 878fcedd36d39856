#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python bench.py 2>gpurun_out/c54.err | tail -1 > gpurun_out/c54_default.json
python - <<PY
import json
r=json.loads(open("gpurun_out/c54_default.json").read())
def show(n, r):
  print(n, round(r["value"],1), round(r["ms_per_step"],2), r["clocks"]["sm_mhz"], r["clocks"].get("power_w"), r["roofline"]["frac"] if r.get("roofline") else None, "e2e", round(r["e2e"]["value"],1), r["gpu_launches"])
show("c4", r)
for k, v in r.get("extra", {}).items():
  if isinstance(v, dict) and "value" in v: show(k, v)
PY
timeout 600 python bench.py --workload c4 --no-extras --no-cpu-baseline --global-batch 64 --steps 5 --warmup 3 2>>gpurun_out/c54.err | tail -1 | python -c "
import json,sys
r=json.loads(sys.stdin.read()); print('c4 gb=64', round(r['value'],1), 'e2e', round(r['e2e']['value'],1), r['config']['execution'][:20])"
timeout 600 python bench.py --workload c3 --no-extras --no-cpu-baseline --global-batch 32 --steps 8 --warmup 3 2>>gpurun_out/c54.err | tail -1 | python -c "
import json,sys
r=json.loads(sys.stdin.read()); print('c3 gb=32', round(r['value'],1), 'e2e', round(r['e2e']['value'],1))"
tail -n 3 gpurun_out/c54.err
