#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/v_pytest.log
for w in c4_1M_1080p_K7 sb_150k_512x288_K9 c4L_1M_1080p_K7; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --workload $w 2> gpurun_out/v_$w.err | tee gpurun_out/v_$w.json | python tools/show_bench.py | sed -n 1,2p
done
MOBGS_FUSED_COUNT=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2> gpurun_out/v_nofuse.err | tee gpurun_out/v_nofuse.json | python tools/show_bench.py | sed -n 1,2p
