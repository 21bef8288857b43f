#!/bin/bash
# (GPU box) graph tests + bench with extras
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_graphs_gpu.py -q --tb=short -p no:cacheprovider ) > gpurun_out/tests_graph.txt 2>&1
tail -25 gpurun_out/tests_graph.txt
timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/bench.err > gpurun_out/bench.json
grep -v Warning gpurun_out/bench.err | tail -5
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    print("ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
    print("cuda_graph", json.dumps(d.get("cuda_graph"), indent=None)[:2500])
except Exception as e:
    print("bench parse failed", e)
PY
