#!/bin/bash
# (GPU box) round-2 session-3 ncu evidence: launch list of the headline step + one --set full capture of each blend kernel
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2c_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu3_launches.log 2>&1
for k in blend_bwd_tr blend_fwd; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r2c_$k \
    python bench.py --no-cpu-baseline --no-extras --steps 2 --warmup 1 > gpurun_out/ncu3_$k.log 2>&1
done
ls -la gpurun_out/r2c_*.ncu-rep gpurun_out/r2c_launches.csv
