#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
MVB_F16F8=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c8_bench_f16f8.json 2> gpurun_out/c8_bench_f16f8.err
MVB_F16F8=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c8_bench_bf16.json 2> gpurun_out/c8_bench_bf16.err
python - <<'PY'
import json
for n in ("f16f8","bf16"):
  try:
    d=json.loads(open("gpurun_out/c8_bench_%s.json"%n).read().strip().splitlines()[-1])
    print(n, d["value"], d["ms_per_step"], d["clocks"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["roofline"]["cell_share_of_step"], d["e2e"]["value"])
  except Exception as e: print(n, "ERR", e)
PY
ncu --set full --clock-control none --import-source on -k regex:cell_fwd_kernel -s 3 -c 1 -o gpurun_out/c8_f16f8 python tools/gpu_probe_cell_time.py 2048 16 > gpurun_out/c8_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cell_fwd_kernel -s 3 -c 1 -o gpurun_out/c8_bf16 python tools/gpu_probe_cell_time.py 2048 2 >> gpurun_out/c8_ncu.log 2>&1
tail -2 gpurun_out/c8_ncu.log
