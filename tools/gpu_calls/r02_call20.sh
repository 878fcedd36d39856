#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ddp_gpu.py -q -x 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/c20_bench_n2.json 2> gpurun_out/c20_bench_n2.err; tail -5 gpurun_out/c20_bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c20_bench_n2.json").read().strip().splitlines()[-1])
def show(n, d):
  print(n, "n_gpus", d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],1), d["roofline"]["frac"], "e2e", round(d["e2e"]["value"],1), d.get("allreduce"))
show("c4", d)
for k, v in d.get("extra", {}).items(): show(k, v); print(v.get("ddp_equivalence"))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-300
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_parity_gpu.py -q -x -k "cell_f16f8_golden and dec_cx32 or fanout_equals" 2>&1 | tail -6
