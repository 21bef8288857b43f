#!/bin/bash
# (GPU box) full GPU test suite on the product build, then bench + ops tests for each prebuilt unit-size variant.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/tests_L16.txt 2>&1; tail -5 gpurun_out/tests_L16.txt
for L in 16 8 32; do
  cp build/variants/lib_L$L.so mobgs_b200/libmobgs_b200.so
  echo "=== L=$L"
  python bench.py --no-cpu-baseline --steps 20 2>gpurun_out/bench_L$L.err | tee gpurun_out/bench_L$L.json | python tools/show_bench.py | head -4
  if [ $L != 16 ]; then ( time python -m pytest tests/test_ops_gpu.py tests/test_golden_gpu.py -m gpu -x -q ) > gpurun_out/tests_L$L.txt 2>&1; tail -3 gpurun_out/tests_L$L.txt; fi
done
cp build/variants/lib_L16.so mobgs_b200/libmobgs_b200.so
