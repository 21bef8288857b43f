#!/bin/bash
# (GPU box) in-kernel rays (templated) + recorded-entries binning: parity tests + A/B bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/k_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/k_new.json 2> gpurun_out/k_new.err; echo "new rc=$?"
MOBGS_RECORD_ENTRIES=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --ray-images > gpurun_out/k_old.json 2> gpurun_out/k_old.err; echo "old rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --workload sb_150k_512x288_K9 > gpurun_out/k_sb.json 2> gpurun_out/k_sb.err; echo "sb rc=$?"
cat gpurun_out/k_new.json gpurun_out/k_old.json gpurun_out/k_sb.json | python tools/show_bench.py | tail -30
timeout 200 python tools/diag_step_gaps.py > gpurun_out/k_diag.json 2> gpurun_out/k_diag.err
