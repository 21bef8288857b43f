#!/bin/bash
# (GPU box) round-2 ncu evidence: launch list of the headline step + one --set full capture per hot kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
for k in blend_bwd blend_fwd tile_rank_sort synth_project_bwd; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r2_$k \
    python bench.py --no-cpu-baseline --no-extras --steps 2 --warmup 1 > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out | tail -12
