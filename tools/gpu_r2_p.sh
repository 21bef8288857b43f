#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_render_gpu.py -m gpu -x -q > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/p_pytest.log
for w in c4L_1M_1080p_K7; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --workload $w 2> gpurun_out/p_$w.err | tee gpurun_out/p_$w.json | python tools/show_bench.py | sed -n 1,2p
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/p_full.json 2> gpurun_out/p_full.err; echo "full rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/p_full.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['host_sync_per_step'])
f=d['full_step']; print({k:f[k] for k in f if k.endswith('_ms')}); print(f.get('full_step_kernel_ms'))
print(d['gpu_on_reference_config']['gpu_ms_per_step'], d['gpu_on_reference_config']['gpu_e2e_ms_per_step'])
print(d.get('strong_scaling'))
P
