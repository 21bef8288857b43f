#!/bin/bash
# (GPU box) a11 evidence: HexPlane + MLP forward / backward timing with the device-side operand pack, and one ncu --set full
# capture of each tcgen05 kernel (tensor-pipe utilisation)
mkdir -p gpurun_out
python tests/perf/bench_hexplane.py > gpurun_out/r2c_bench_hexplane.json 2>gpurun_out/r2c_bench_hexplane.err
cat gpurun_out/r2c_bench_hexplane.json
for k in hexplane_mlp_fwd hexplane_mlp_bwd hexplane_wgrad; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r2c_$k \
    python tests/perf/bench_hexplane.py > gpurun_out/ncu3_$k.log 2>&1
done
ls -la gpurun_out/r2c_hexplane*.ncu-rep
