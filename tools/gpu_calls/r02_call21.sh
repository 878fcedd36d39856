#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
for gb in 64 128; do
timeout 300 python bench.py --workload c4 --global-batch $gb --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4 gb=$gb', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), round(d['ms_per_step'],2))"
done
