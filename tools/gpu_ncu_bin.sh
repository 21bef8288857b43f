#!/bin/bash
mkdir -p gpurun_out
for kn in tile_count tile_emit; do
  ncu --set full --clock-control none --import-source on -k regex:${kn}_kernel -s 2 -c 1 -f -o gpurun_out/${kn}_c4 \
    python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/ncu_$kn.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:${kn}_kernel -s 2 -c 1 -f -o gpurun_out/${kn}_sb \
    python bench.py --no-cpu-baseline --steps 2 --warmup 1 --workload sb_150k_512x288_K9 > gpurun_out/ncu_${kn}_sb.log 2>&1
done
ls gpurun_out/*.ncu-rep
