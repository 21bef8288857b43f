#!/bin/bash
# (GPU box) numbers for DESIGN.md's results table: other workloads + the reference-call-pattern baseline.
mkdir -p gpurun_out
for w in c2_200k_960x540_K7 c3_500k_960x540_K7 c4_1M_1080p_K9 sb_150k_512x288_K9; do
  python bench.py --no-cpu-baseline --steps 20 --workload $w 2>/dev/null | tee gpurun_out/bench_$w.json | python tools/show_bench.py | head -2
done
for w in c4_1M_1080p_K7 c2_200k_960x540_K7 sb_150k_512x288_K9; do
  short=c4; [ $w = c2_200k_960x540_K7 ] && short=c2; [ $w = sb_150k_512x288_K9 ] && short=sb
  python tests/perf/bench_unfused.py $w 5 2>/dev/null | tail -1 | tee gpurun_out/unfused_baseline_$short.json
done
