#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_flow_warp_gpu.py -m gpu -x -q 2>&1 | tail -12
python tests/perf/bench_flow_warp.py | tee gpurun_out/flow_warp.json
