#!/bin/bash
# (GPU box) numbers for DESIGN.md: headline line with every extra + the other workloads
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/t_c4.json 2> gpurun_out/t_c4.err; echo "c4 rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/t_ref.json 2> gpurun_out/t_ref.err; echo "ref rc=$?"
for w in c2_200k_960x540_K7 c3_500k_960x540_K7 c4_1M_1080p_K9 sb_150k_512x288_K9 c4L_1M_1080p_K7; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/t_$w.json 2> gpurun_out/t_$w.err; echo "$w rc=$?"
done
timeout 200 python tests/perf/bench_hexplane.py > gpurun_out/t_hexplane.json 2>&1
