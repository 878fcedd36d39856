#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 300 python tools/gpu_probe_d2h.py 2>&1 | tail -8
