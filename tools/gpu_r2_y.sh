#!/bin/bash
# (GPU box) ncu captures of the loss kernels (f1, f2), which round 1 only timed
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for k in photo_loss_fwd photo_loss_bwd; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o gpurun_out/r2b_$k python tests/perf/bench_loss.py > gpurun_out/ncu2_$k.log 2>&1
done
for k in flow_warp_fwd flow_warp_bwd; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o gpurun_out/r2b_$k python tests/perf/bench_flow_warp.py > gpurun_out/ncu2_$k.log 2>&1
done
ls -la gpurun_out/r2b_photo* gpurun_out/r2b_flow* 2>&1 | tail -5
timeout 200 python tests/perf/bench_loss.py > gpurun_out/y_loss.json 2>&1; tail -2 gpurun_out/y_loss.json | cut -c1-600
timeout 200 python tests/perf/bench_flow_warp.py > gpurun_out/y_flow.json 2>&1; tail -2 gpurun_out/y_flow.json | cut -c1-600
