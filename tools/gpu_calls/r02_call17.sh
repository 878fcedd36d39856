#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x 2>&1 | tail -4
bash tools/r02_call16.sh
