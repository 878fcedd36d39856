#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for m in 2 1; do
MVB_CELL_PAIR=$m timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c10_bench_c4_pair$m.json 2> gpurun_out/c10_bench_c4_pair$m.err
done
timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c10_bench_c3.json 2> gpurun_out/c10_bench_c3.err
python - <<'PY'
import json
for n in ("c4_pair2","c4_pair1","c3"):
  try:
    d=json.loads(open("gpurun_out/c10_bench_%s.json"%n).read().strip().splitlines()[-1])
    print(n, d["value"], d["ms_per_step"], d["clocks"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["roofline"]["cell_share_of_step"], d["e2e"]["value"])
  except Exception as e: print(n, "ERR", e)
PY
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
