#!/bin/bash
# (GPU box) one ncu --set full capture of the fused blend backward kernel of the headline step.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:blend_bwd -s 2 -c 1 -f -o gpurun_out/blend_bwd \
  python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/ncu_bwd.log 2>&1
tail -2 gpurun_out/ncu_bwd.log
