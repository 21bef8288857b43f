#!/bin/bash
mkdir -p gpurun_out
python tests/perf/bench_loss.py | tee gpurun_out/photo_loss.json
