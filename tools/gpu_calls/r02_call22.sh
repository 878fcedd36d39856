#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 python -m pytest tests/test_train_gpu.py -q -k "optimizers" 2>&1 | tail -3
timeout 300 python tools/gpu_probe_e2e.py 64 2>&1 | tail -6
