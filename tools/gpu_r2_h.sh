#!/bin/bash
# 8 GPUs: weak scaling with / without the overlapped all-reduce (+ strong scaling extra)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/h_n$N.json 2> gpurun_out/h_n$N.err
echo "rc=$?"; tail -2 gpurun_out/h_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-overlap > gpurun_out/h_n${N}_noov.json 2> gpurun_out/h_n${N}_noov.err
echo "rc=$?"; tail -2 gpurun_out/h_n${N}_noov.err
