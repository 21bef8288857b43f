#!/bin/bash
# round 2, call A: GPU tests + new bench contract + sensitivity workload
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -15 gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/a_bench.json; tail -5 gpurun_out/a_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/a_ref.json 2> gpurun_out/a_ref.err; echo "ref rc=$?"
timeout 400 python bench.py --workload c4L_1M_1080p_K7 --steps 5 --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/a_c4L.json 2> gpurun_out/a_c4L.err; echo "c4L rc=$?"
tail -c 1500 gpurun_out/a_c4L.json; tail -5 gpurun_out/a_c4L.err
