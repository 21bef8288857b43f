#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_losses_gpu.py -m gpu -x -q 2>&1 | tail -3
python tests/perf/bench_loss.py | tee gpurun_out/photo_loss.json
python bench.py --no-cpu-baseline --steps 20 | tee gpurun_out/bench.json | python tools/show_bench.py | head -5
