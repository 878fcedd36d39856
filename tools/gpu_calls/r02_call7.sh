#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x 2>&1 | tail -15 > gpurun_out/c7_tests.log; cat gpurun_out/c7_tests.log
L=gpurun_out/c7_probe.log; : > $L
timeout 120 python tools/gpu_probe_cell_time.py 2048 2 16 1 3 >> $L 2>&1
MVB_CELL_ABL=7 timeout 120 python tools/gpu_probe_cell_time.py 2048 2 16 >> $L 2>&1
MVB_CELL_ABL=4 timeout 120 python tools/gpu_probe_cell_time.py 2048 2 16 >> $L 2>&1
MVB_CELL_MULTICAST=0 timeout 120 python tools/gpu_probe_cell_time.py 2048 2 16 >> $L 2>&1
cat $L
