#!/bin/bash
# (GPU box, 2 GPUs) gradient all-reduce overlap with a high-priority comm stream: plain vs --overlap (NCCL stream priority on/off)
mkdir -p gpurun_out
run() {  # name, env, flags
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 30 --warmup 5 --no-extras $3 2>gpurun_out/n2_$1.err > gpurun_out/n2_$1.json
  python - "$1" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/n2_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4))
PY
}
run plain A=1 ""
run overlap A=1 "--overlap"
run overlap_hp TORCH_NCCL_HIGH_PRIORITY=1 "--overlap"
run plain_hp TORCH_NCCL_HIGH_PRIORITY=1 ""
