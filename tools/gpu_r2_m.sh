#!/bin/bash
# (GPU box) bucketed rank sort: list tests + bench (c4 and sb)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_render_gpu.py tests/test_golden_gpu.py tests/test_fullsize_properties_gpu.py -m gpu -x -q > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/m_pytest.log
for w in c4_1M_1080p_K7 sb_150k_512x288_K9 c4L_1M_1080p_K7; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --workload $w 2> gpurun_out/m_$w.err | tee gpurun_out/m_$w.json | python tools/show_bench.py | sed -n 1,2p
done
timeout 200 python tools/diag_step_gaps.py > gpurun_out/m_diag.json 2> gpurun_out/m_diag.err
python - <<'P'
import json
d=json.load(open('gpurun_out/m_diag.json'))
for r in d['device_activities_ms_per_step'][:14]: print(r)
P
