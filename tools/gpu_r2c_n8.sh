#!/bin/bash
# (GPU box, 8 GPUs) weak scaling at N = 8: one multimem all-reduce after the backward vs the overlapped, high-priority one
mkdir -p gpurun_out
run() {  # name, flags
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 30 --warmup 5 $2 2>gpurun_out/n8_$1.err > gpurun_out/n8_$1.json
  python - "$1" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/n8_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), d.get("gradient_allreduce"), "strong", (d.get("strong_scaling") or {}).get("ms_per_step"))
PY
}
run plain "--no-overlap --no-extras"
run overlap ""
