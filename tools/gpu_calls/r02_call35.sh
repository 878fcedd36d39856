#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c4 --global-batch 128 --no-extras --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/c35_n2.out 2> gpurun_out/c35_n2.err
echo rc=$?
grep "^{" gpurun_out/c35_n2.out | tail -1 > gpurun_out/c35_n2.json
python - <<PY
import json
r=json.loads(open("gpurun_out/c35_n2.json").read())
print("c4 N=2 gb=128", round(r["value"],1), round(r["ms_per_step"],2), r["clocks"]["sm_mhz"], "e2e", round(r["e2e"]["value"],1), r["config"]["execution"][:30])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload c3 --global-batch 64 --no-extras --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/c35_n2c3.out 2> gpurun_out/c35_n2c3.err
echo rc=$?
grep "^{" gpurun_out/c35_n2c3.out | tail -1 > gpurun_out/c35_n2c3.json
python - <<PY
import json
r=json.loads(open("gpurun_out/c35_n2c3.json").read())
print("c3 N=2 gb=64", round(r["value"],1), round(r["ms_per_step"],2), r["clocks"]["sm_mhz"], "e2e", round(r["e2e"]["value"],1), r["config"]["l2"][:40])
PY
tail -3 gpurun_out/c35_n2.err gpurun_out/c35_n2c3.err
