#!/bin/bash
# (GPU box) round-2 (second session) ncu evidence: launch list of the headline step + one --set full capture per hot kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu2_launches.log 2>&1
for k in blend_bwd_tr blend_fwd tile_count_entries tile_bucket_sort synth_project_bwd; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r2b_$k \
    python bench.py --no-cpu-baseline --no-extras --steps 2 --warmup 1 > gpurun_out/ncu2_$k.log 2>&1
done
# the large-segment sort only runs on the heavy-footprint workload
ncu --set full --clock-control none --import-source on -k regex:tile_msd_sort -s 1 -c 1 -f -o gpurun_out/r2b_tile_msd_sort_c4L \
  python bench.py --no-cpu-baseline --no-extras --steps 1 --warmup 1 --workload c4L_1M_1080p_K7 > gpurun_out/ncu2_msd.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8
