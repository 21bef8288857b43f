#!/bin/bash
# (GPU box) pipelined HexPlane forward: tests + bench (under a short timeout: a wrong barrier traps, it must not hang)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_hexplane_gpu.py -m gpu -x -q > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/u_pytest.log
timeout 200 python tests/perf/bench_hexplane.py > gpurun_out/u_hex_pipe.json 2>&1; tail -1 gpurun_out/u_hex_pipe.json | cut -c1-700
MOBGS_HEX_PIPELINED=0 timeout 200 python tests/perf/bench_hexplane.py > gpurun_out/u_hex_serial.json 2>&1; tail -1 gpurun_out/u_hex_serial.json | cut -c1-400
