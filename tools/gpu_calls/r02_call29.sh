#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gnn_rows -s 3 -c 1 -f -o gpurun_out/r02_gnn_rows python tools/gpu_probe_gnn_head.py 4096 2>&1 | tail -3
ls -la gpurun_out/*.ncu-rep
