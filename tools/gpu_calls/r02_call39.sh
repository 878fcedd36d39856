#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/c39_n8.out 2> gpurun_out/c39_n8.err
echo rc=$?
grep "^{" gpurun_out/c39_n8.out | tail -1 > gpurun_out/c39_n8.json
python - <<PY
import json
r=json.loads(open("gpurun_out/c39_n8.json").read())
def show(n, r):
  print(n, round(r["value"],1), round(r["ms_per_step"],2), r["clocks"]["sm_mhz"], "e2e", round(r["e2e"]["value"],1))
show("c4", r)
for k, v in r.get("extra", {}).items():
  if isinstance(v, dict) and "value" in v:
    show(k, v)
    if v.get("allreduce"): print(v["allreduce"]); print(v.get("ddp_equivalence"))
PY
tail -n 5 gpurun_out/c39_n8.err
