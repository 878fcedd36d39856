#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_dropin_gpu.py tests/test_parity_gpu.py -q -k "pinned or gnn" 2>&1 | tail -8
MVB_GNN_ROWS=0 timeout 900 python -m pytest tests/test_parity_gpu.py -q -k "gnn" 2>&1 | tail -3
timeout 300 python tools/gpu_probe_gnn_head.py 2>&1 | tail -6
MVB_GNN_ROWS=0 timeout 300 python tools/gpu_probe_gnn_head.py 2>&1 | head -2
