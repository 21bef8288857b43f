#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_render_gpu.py tests/test_golden_gpu.py tests/test_ops_gpu.py tests/test_train_loop.py -m gpu -x -q > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/z_pytest.log
bash tools/gpu_r2_r.sh product nowgt
