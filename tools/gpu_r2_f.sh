#!/bin/bash
# 2 GPUs: tests + weak scaling with / without the overlapped all-reduce
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_flatgrad_gpu.py tests/test_hexplane_gpu.py tests/test_densify.py -m gpu -q 2>&1 | tail -5
for flag in "" "--no-overlap"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras $flag > gpurun_out/f_n2$flag.json 2> gpurun_out/f_n2$flag.err
  echo "rc=$?"; tail -c 700 gpurun_out/f_n2$flag.json | head -c 700; echo; tail -3 gpurun_out/f_n2$flag.err
done
