#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
M=gpu__time_duration.sum,gpc__cycles_elapsed.avg.per_second,sm__cycles_elapsed.avg,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
: > gpurun_out/c16_ncu.csv
for cfg in "28 16" "28 2" "29 16" "4 16" "0 16" "0 2"; do set -- $cfg
  echo "== abl $1 planes $2" >> gpurun_out/c16_ncu.csv
  MVB_CELL_ABL=$1 timeout 200 ncu --metrics $M --clock-control none -k regex:cell_fwd_kernel -s 3 -c 1 --csv python tools/gpu_probe_cell_time.py 4096 $2 2>/dev/null | grep -E "cell_fwd_kernel" | awk -F'","' '{print $(NF-2)","$(NF-1)","$NF}' >> gpurun_out/c16_ncu.csv
done
cat gpurun_out/c16_ncu.csv
