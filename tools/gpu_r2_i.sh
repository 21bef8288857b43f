#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/i_n$N.json 2> gpurun_out/i_n$N.err
echo "rc=$?"; grep -v "OMP_NUM\|\*\*\*\*\|^$" gpurun_out/i_n$N.err | tail -4
