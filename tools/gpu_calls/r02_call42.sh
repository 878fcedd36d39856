#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
echo "== racecheck gnn"; timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_parity_gpu.py -q -x -k "gnn" 2>&1 | grep -v "^$" | tail -12
echo "== memcheck gnn + simaug + metrics"; timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_parity_gpu.py tests/test_train_gpu.py -q -x -k "gnn or multiview or eval_metrics" 2>&1 | grep -v "^$" | tail -8
echo "== synccheck gnn"; timeout 900 compute-sanitizer --tool synccheck python -m pytest tests/test_parity_gpu.py -q -x -k "gnn_shapes" 2>&1 | grep -v "^$" | tail -6
