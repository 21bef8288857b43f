#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -5 gpurun_out/c_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/c_bench.err
# launch list of one headline step (every kernel incl. torch's): what is the "everything else"?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c_launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/c_ncu_bench.log 2>&1; echo "ncu rc=$?"
