#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -q -x 2>&1 | tail -8 > gpurun_out/c9_tests.log; cat gpurun_out/c9_tests.log
L=gpurun_out/c9_probe.log; : > $L
timeout 120 python tools/gpu_probe_cell_time.py 2048 2 16 1 >> $L 2>&1
MVB_CELL_PAIR=1 timeout 120 python tools/gpu_probe_cell_time.py 2048 2 16 >> $L 2>&1
MVB_CELL_ABL=7 timeout 120 python tools/gpu_probe_cell_time.py 2048 2 16 >> $L 2>&1
MVB_CELL_ABL=4 timeout 120 python tools/gpu_probe_cell_time.py 2048 2 16 >> $L 2>&1
cat $L
