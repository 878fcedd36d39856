#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 python -m pytest tests -m gpu -q -k "atsize or simaug or fanout" -s 2>&1 | grep -v "^$" | tail -14 | cut -c1-400
bash tools/r02_call14.sh
