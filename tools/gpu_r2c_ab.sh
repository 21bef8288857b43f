#!/bin/bash
# (GPU box) A/B of prebuilt library variants build/variants/lib_<name>.so: headline kernel table (+ parity tests when TESTS=1)
mkdir -p gpurun_out
cp mobgs_b200/libmobgs_b200.so /tmp/lib_keep.so
for v in "$@"; do
  cp build/variants/lib_$v.so mobgs_b200/libmobgs_b200.so
  echo "=== variant $v"
  python bench.py --no-cpu-baseline --no-extras --steps 30 --warmup 5 2>gpurun_out/ab_$v.err > gpurun_out/ab_$v.json
  python - "$v" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/ab_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), {k.replace("mobgs_", ""): round(v, 4) for k, v in d["kernel_ms_per_step"].items()})
PY
  [ "$TESTS" = 1 ] && python -m pytest tests/test_ops_gpu.py tests/test_golden_gpu.py tests/test_render_gpu.py tests/test_train_loop.py tests/test_fullsize_parity_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
done
cp /tmp/lib_keep.so mobgs_b200/libmobgs_b200.so
