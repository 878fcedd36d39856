#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1200 python -m pytest tests/test_train_gpu.py -q -x -s -k "simaug or mixup" 2>&1 | grep -v "^$" | tail -22
