#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 300 python -m pytest tests/test_flatgrad_gpu.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed" | cut -c1-250 | head -12
