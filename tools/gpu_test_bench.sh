#!/bin/bash
# (GPU box) GPU test suite + one headline bench line of the current build.
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/tests.txt 2>&1; tail -15 gpurun_out/tests.txt
python bench.py --no-cpu-baseline --steps 20 2>gpurun_out/bench.err | tee gpurun_out/bench.json | python tools/show_bench.py | head -4
