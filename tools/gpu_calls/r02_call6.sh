#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
L=gpurun_out/c6_probe.log; : > $L
MVB_CELL_ABL=7 timeout 120 python tools/gpu_probe_cell_time.py 2048 2 1 3 16 >> $L 2>&1
MVB_CELL_ABL=7 MVB_CELL_MULTICAST=0 timeout 120 python tools/gpu_probe_cell_time.py 2048 2 16 >> $L 2>&1
cat $L
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/c6_tests.log; cat gpurun_out/c6_tests.log
