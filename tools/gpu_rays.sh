#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_cameras_gpu.py -m gpu -x -q 2>&1 | tail -5
bash tools/gpu_test_bench.sh
