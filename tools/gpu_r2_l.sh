#!/bin/bash
# (GPU box) full GPU tests on the product build, then A/B of the backward variants, then the product bench with all keys
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cp mobgs_b200/libmobgs_b200.so /tmp/lib_product.so
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/l_pytest.log
for v in base leanA mom both; do
  cp build/variants/lib_$v.so mobgs_b200/libmobgs_b200.so
  echo "=== variant $v"
  timeout 200 python bench.py --no-cpu-baseline --no-extras --steps 20 --warmup 5 2>gpurun_out/l_$v.err | tee gpurun_out/l_$v.json | python tools/show_bench.py | sed -n 1,2p
done
cp /tmp/lib_product.so mobgs_b200/libmobgs_b200.so
python - <<'P'
import json
d=json.load(open('gpurun_out/l_both.json'))
print({k:d[k] for k in ('ms_per_step','host_sync_per_step')}, d['e2e']['ms_per_step'])
P
