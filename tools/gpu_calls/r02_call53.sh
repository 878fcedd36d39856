#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
