#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_dropin_gpu.py tests/test_parity_gpu.py -q -x 2>&1 | tail -5
echo "== graph auto"; timeout 300 python tools/gpu_probe_e2e.py 64 2>&1 | tail -8
echo "== eager";      MVB_CUDA_GRAPH=0 timeout 300 python tools/gpu_probe_e2e.py 64 2>&1 | tail -8
timeout 300 python tools/gpu_probe_gnn_head.py 2>&1 | tail -6
timeout 600 python bench.py --workload c4 --no-extras --steps 3 --warmup 2 2>gpurun_out/c26.err | tail -1 > gpurun_out/c26_c4.json
python - <<'PY'
import json
r=json.loads(open("gpurun_out/c26_c4.json").read())
print("c4", r["value"], r["ms_per_step"], r["clocks"]["sm_mhz"], r["roofline"]["frac"], "e2e", r["e2e"]["value"])
PY
