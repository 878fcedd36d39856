#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
echo "== memcheck new cell paths + mixup"; timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_parity_gpu.py tests/test_train_gpu.py -q -x -k "xsparse or xdense or mixup or forward_graph" 2>&1 | grep -v "^$" | tail -6
echo "== racecheck xsparse table"; timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_parity_gpu.py -q -x -k "xsparse" 2>&1 | grep -v "^$" | tail -4
