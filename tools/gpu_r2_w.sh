#!/bin/bash
# (GPU box, N GPUs) weak-scaling bench line with the final build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/w_n$N.json 2> gpurun_out/w_n$N.err
echo "rc=$?"; grep -v "OMP_NUM\|\*\*\*\*\|^$" gpurun_out/w_n$N.err | tail -4
python - <<P
import json
d=json.loads(open('gpurun_out/w_n$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d.get('gradient_allreduce'), d.get('strong_scaling'))
P
