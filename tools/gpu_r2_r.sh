#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cp mobgs_b200/libmobgs_b200.so /tmp/lib_product.so
for v in "$@"; do
  cp build/variants/lib_$v.so mobgs_b200/libmobgs_b200.so
  echo "=== variant $v"
  timeout 200 python bench.py --no-cpu-baseline --no-extras --steps 20 --warmup 5 2>gpurun_out/r_$v.err | tee gpurun_out/r_$v.json | python tools/show_bench.py | sed -n 1,2p
done
cp /tmp/lib_product.so mobgs_b200/libmobgs_b200.so
