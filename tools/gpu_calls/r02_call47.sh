#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -s -k "cell or rollout or forward_graph" 2>&1 | grep -v "^$" | tail -12
for spec in "c4 512 3" "c3 256 8"; do
set -- $spec
timeout 600 python bench.py --workload $1 --no-extras --no-cpu-baseline --global-batch $2 --steps $3 --warmup 3 2>>gpurun_out/c47.err | tail -1 > gpurun_out/c47_$1_$2.json
python - <<PY
import json
r=json.loads(open("gpurun_out/c47_$1_$2.json").read())
print("$1 gb=$2", round(r["value"],1), round(r["ms_per_step"],2), r["clocks"]["sm_mhz"], r["clocks"].get("power_w"), r["roofline"]["frac"], "e2e", round(r["e2e"]["value"],1))
PY
MVB_REG_XDENSE=0 timeout 600 python bench.py --workload $1 --no-extras --no-cpu-baseline --global-batch $2 --steps $3 --warmup 3 2>>gpurun_out/c47.err | tail -1 > gpurun_out/c47_$1_$2_off.json
python - <<PY
import json
r=json.loads(open("gpurun_out/c47_$1_$2_off.json").read())
print("$1 gb=$2 (xdense off)", round(r["value"],1), round(r["ms_per_step"],2), r["clocks"]["sm_mhz"], r["clocks"].get("power_w"), r["roofline"]["frac"], "e2e", round(r["e2e"]["value"],1))
PY
done
