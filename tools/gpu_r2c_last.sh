#!/bin/bash
# (GPU box) last sanity of the committed tree: GPU suite, smoke, default-flag bench line (without the 12 s CPU sample)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/r2c_tests_last.txt 2>&1
grep "passed\|failed" gpurun_out/r2c_tests_last.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/r2c_last.err > gpurun_out/r2c_last.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c_last.json").read().strip().splitlines()[-1])
print("ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), "traffic", d["roofline"]["traffic"], d["roofline"]["traffic_source"],
      "frac", round(d["roofline"]["frac"], 4), "full", round(d["full_step"]["full_step_ms"], 2), "graph", round(d["cuda_graph"]["ms_per_step"], 3),
      round(d["cuda_graph"]["sb_150k_512x288_K9"]["ms_per_step"], 3), "launches", d["gpu_launches"], d["steps"])
PY
