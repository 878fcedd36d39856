#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1200 python -m pytest tests/test_train_gpu.py -q -s -k "whole_model or against_reference_execution" 2>&1 | grep "worst\|passed\|failed" | cut -c1-700
