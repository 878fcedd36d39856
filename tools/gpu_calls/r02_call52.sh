#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -s -k "xsparse or rollout or forward_graph or full_size" 2>&1 | grep -v "^$\|atsize" | tail -8
for spec in "c3 256 8" "c4 512 3"; do
set -- $spec
timeout 600 python bench.py --workload $1 --no-extras --no-cpu-baseline --global-batch $2 --steps $3 --warmup 3 2>>gpurun_out/c52.err | tail -1 > gpurun_out/c52_$1_$2.json
python - <<PY
import json
r=json.loads(open("gpurun_out/c52_$1_$2.json").read())
print("$1 gb=$2", round(r["value"],1), round(r["ms_per_step"],2), r["clocks"]["sm_mhz"], r["clocks"].get("power_w"), r["roofline"]["frac"], "e2e", round(r["e2e"]["value"],1))
PY
MVB_ENC_XSPARSE=0 timeout 600 python bench.py --workload $1 --no-extras --no-cpu-baseline --global-batch $2 --steps $3 --warmup 3 2>>gpurun_out/c52.err | tail -1 > gpurun_out/c52_$1_$2_off.json
python - <<PY
import json
r=json.loads(open("gpurun_out/c52_$1_$2_off.json").read())
print("$1 gb=$2 (xsparse off)", round(r["value"],1), round(r["ms_per_step"],2), r["clocks"]["sm_mhz"], r["clocks"].get("power_w"), r["roofline"]["frac"], "e2e", round(r["e2e"]["value"],1))
PY
done
