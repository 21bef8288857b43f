#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_ops_gpu.py -x -q -k "speculative or recorded_entries" > gpurun_out/san.log 2>&1; echo "rc=$?"
grep -n "Invalid\|at 0x\|by thread\|Address\|=========     at\|in .*kernel" gpurun_out/san.log | head -30; tail -5 gpurun_out/san.log
