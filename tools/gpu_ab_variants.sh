#!/bin/bash
# (GPU box) bench (+ the render/ops/golden tests when TESTS=1) for each prebuilt library variant build/variants/lib_<name>.so
mkdir -p gpurun_out
for v in "$@"; do
  cp build/variants/lib_$v.so mobgs_b200/libmobgs_b200.so
  echo "=== variant $v"
  python bench.py --no-cpu-baseline --steps 20 2>gpurun_out/bench_$v.err | tee gpurun_out/bench_$v.json | python tools/show_bench.py | sed -n 1,2p
  [ "$TESTS" = 1 ] && python -m pytest tests/test_ops_gpu.py tests/test_golden_gpu.py tests/test_render_gpu.py -m gpu -x -q 2>&1 | tail -1
done
