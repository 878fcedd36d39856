#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_train_gpu.py -q -x -k "simaug" 2>&1 | tail -25
