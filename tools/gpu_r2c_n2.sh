#!/bin/bash
# (GPU box, 2 GPUs) weak scaling sanity of the current build
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 20 --warmup 5 --no-extras 2>gpurun_out/bench_n2.err > gpurun_out/bench_n2.json
grep -v Warning gpurun_out/bench_n2.err | tail -5
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print("n", d["n_gpus"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d.get("gradient_allreduce"))
PY
