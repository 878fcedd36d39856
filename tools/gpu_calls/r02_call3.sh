#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
L=gpurun_out/c3_probe.log; : > $L
python tools/gpu_probe_cell_time.py 2048 2 16 1 >> $L 2>&1
MVB_CELL_MULTICAST=0 python tools/gpu_probe_cell_time.py 2048 2 16 >> $L 2>&1
for a in 1 2 4 5 6 7; do MVB_CELL_ABL=$a python tools/gpu_probe_cell_time.py 2048 16 >> $L 2>&1; done
MVB_CELL_ABL=4 python tools/gpu_probe_cell_time.py 2048 2 >> $L 2>&1
cat $L
ncu --set full --clock-control none --import-source on -k regex:cell_fwd_kernel -s 3 -c 1 -o gpurun_out/c3_f16f8 python tools/gpu_probe_cell_time.py 2048 16 > gpurun_out/c3_ncu.log 2>&1
tail -3 gpurun_out/c3_ncu.log
