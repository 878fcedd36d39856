#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi topo -m 2>&1 | head -8
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH,ENV timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/c32_n2.out 2> gpurun_out/c32_n2.err
tail -1 gpurun_out/c32_n2.out > gpurun_out/c32_n2.json
python - <<PY
import json
r=json.loads(open("gpurun_out/c32_n2.json").read())
def show(n, r):
  print(n, round(r["value"],1), round(r["ms_per_step"],2), r["clocks"]["sm_mhz"], "e2e", round(r["e2e"]["value"],1))
show("c4", r)
for k, v in r.get("extra", {}).items():
  if isinstance(v, dict) and "value" in v: show(k, v); print({kk: vv for kk, vv in v.items() if kk in ("allreduce",)})
  else: print(k, v)
PY
grep -i "NCCL INFO" gpurun_out/c32_n2.err | grep -i "via\|P2P\|SHM\|NVLS\|channel\|Connected\|transport\|NET/\|topo" | head -30
timeout 600 python -m pytest tests/test_ddp_gpu.py -q 2>&1 | tail -3
