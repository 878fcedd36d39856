#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cell_fwd_kernel -s 13 -c 1 -f -o gpurun_out/r02_cell_final python bench.py --workload c4 --no-extras --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/c38a.log 2>&1
ls -la gpurun_out/r02_cell_final.ncu-rep
