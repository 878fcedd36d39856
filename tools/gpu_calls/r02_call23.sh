#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
echo "== graph auto"; timeout 300 python tools/gpu_probe_e2e.py 64 2>&1 | tail -8
echo "== eager";      MVB_CUDA_GRAPH=0 timeout 300 python tools/gpu_probe_e2e.py 64 2>&1 | tail -8
echo "== graph, small fetch"; PROBE_SMALL_FETCH=1 timeout 300 python tools/gpu_probe_e2e.py 64 2>&1 | tail -8
echo "== eager, small fetch"; PROBE_SMALL_FETCH=1 MVB_CUDA_GRAPH=0 timeout 300 python tools/gpu_probe_e2e.py 64 2>&1 | tail -8
