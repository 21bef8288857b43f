#!/bin/bash
# (GPU box) new full-size binning test, ncu evidence, round-2 ablation
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fullsize_properties_gpu.py tests/test_ray_pose_gpu.py -m gpu -x -q > gpurun_out/t2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t2_pytest.log
bash tools/gpu_r2_ncu2.sh > gpurun_out/t2_ncu.log 2>&1; tail -3 gpurun_out/t2_ncu.log
bash tools/ablation_r2.sh > gpurun_out/r2b_ablation.txt 2>&1; cat gpurun_out/r2b_ablation.txt
