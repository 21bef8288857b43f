#!/bin/bash
# (GPU box, 8 GPUs) overlapped gradient all-reduce with multimem slices vs one multimem all-reduce after the backward
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
for mode in overlap plain; do
  flag=""; [ $mode = overlap ] && flag="--overlap"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 --no-extras $flag > gpurun_out/x_${mode}_n$N.json 2> gpurun_out/x_${mode}_n$N.err
  echo "$mode rc=$?"; grep -v "OMP_NUM\|\*\*\*\*\|^$" gpurun_out/x_${mode}_n$N.err | tail -3
  python - <<P
import json
d=json.loads(open('gpurun_out/x_${mode}_n$N.json').read().strip().splitlines()[-1])
print('$mode', d['n_gpus'], d['ms_per_step'], d['e2e']['ms_per_step'], d.get('gradient_allreduce'))
P
done
