#!/bin/bash
# (GPU box) in-kernel rays: parity tests + A/B bench against the ray-image path
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/j_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/j_pose.json 2> gpurun_out/j_pose.err; echo "pose rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --ray-images > gpurun_out/j_img.json 2> gpurun_out/j_img.err; echo "img rc=$?"
cat gpurun_out/j_pose.json gpurun_out/j_img.json | python tools/show_bench.py | tail -30
