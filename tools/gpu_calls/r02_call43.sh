#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -k "head or gnn or rollout or smoke" 2>&1 | tail -3
timeout 300 python tools/gpu_probe_gnn_head.py 2>&1 | tail -5
