#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 tests/perf/bench_allreduce.py 2>gpurun_out/ar_$1.err | tail -1; }
run 29520
NCCL_ALGO=NVLS run 29521
NCCL_ALGO=Ring run 29522
NCCL_NVLS_ENABLE=0 run 29523
