#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_adam_gpu.py -m gpu -x -q -s 2>&1 | tail -15
python tests/perf/bench_adam.py | tee gpurun_out/adam_1M.json
ncu --set full --clock-control none --import-source on -k regex:adam_kernel -s 3 -c 1 -f -o gpurun_out/fused_adam \
  python tests/perf/bench_adam.py > gpurun_out/ncu_adam.log 2>&1
# the rcp change in blend bwd: full GPU suite + bench
bash tools/gpu_test_bench.sh
