#!/bin/bash
# (GPU box) full GPU test suite, then the numbers for DESIGN.md / profiles (tools/gpu_r2_tables.sh)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s_pytest.log
bash tools/gpu_r2_tables.sh
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/s_smoke.log
