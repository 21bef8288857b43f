#!/bin/bash
# (GPU box) round-2 session 3: full GPU suite + headline bench with all extras (full_step, cuda_graph, strong_scaling).
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/tests.txt 2>&1
tail -40 gpurun_out/tests.txt
timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/bench.err > gpurun_out/bench.json
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    print("ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "kernels", {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()})
    fs = d.get("full_step", {})
    print("full_step", {k: fs.get(k) for k in ("full_step_ms", "render_blurry_view_fwd_bwd_ms", "get_flow_batched_fwd_bwd_ms", "error")})
    print("full_step kernels", fs.get("full_step_kernel_ms"))
    print("cuda_graph", json.dumps(d.get("cuda_graph"), indent=None)[:1500])
    print("strong", d.get("strong_scaling", {}).get("ms_per_step"))
except Exception as e:
    print("bench parse failed", e)
PY
