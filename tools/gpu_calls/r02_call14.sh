#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
L=gpurun_out/c14_probe.log; : > $L
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 100 > gpurun_out/c14_clocks.csv &
SMI=$!
for a in 0 3 4 5 6 7; do echo "== abl $a $(date +%s.%N)" >> $L; MVB_CELL_ABL=$a timeout 120 python tools/gpu_probe_cell_time.py 4096 16 >> $L 2>&1; echo "== end $(date +%s.%N)" >> $L; done
for a in 0 3 4; do echo "== bf16 abl $a" >> $L; MVB_CELL_ABL=$a timeout 120 python tools/gpu_probe_cell_time.py 4096 2 >> $L 2>&1; done
kill $SMI
cat $L
awk -F, '{print $1}' gpurun_out/c14_clocks.csv | sort | uniq -c | sort -rn | head -12
