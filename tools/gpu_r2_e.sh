#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_hexplane_gpu.py -m gpu -q 2>&1 | tail -15
timeout 200 python tests/perf/bench_hexplane.py 2>&1 | tail -12
