#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for v in "A MVB_BENCH_DENSE_FEEDS=0" "B MVB_BENCH_DENSE_FEEDS=1" "C MVB_CUDA_GRAPH=0" "D MVB_BENCH_DENSE_FEEDS=0"; do set -- $v
  env $2 timeout 300 python bench.py --workload c3 --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['e2e']['ms_per_step'])"
done
