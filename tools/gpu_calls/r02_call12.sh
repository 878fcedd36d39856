#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -25
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err; tail -3 gpurun_out/c12_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c12_bench.json").read().strip().splitlines()[-1])
def show(n, d):
  print(n, round(d["value"],1), round(d["ms_per_step"],1), d["clocks"], d["roofline"]["frac"], "e2e", round(d["e2e"]["value"],1), d.get("allreduce"), d.get("cpu_baseline") and d["cpu_baseline"]["value"])
show("c4", d)
for k, v in d.get("extra", {}).items(): show(k, v); print(v.get("ddp_equivalence"))
PY
