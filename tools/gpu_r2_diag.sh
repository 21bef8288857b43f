#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/diag_step_gaps.py > gpurun_out/diag_c4.json 2> gpurun_out/diag_c4.err; echo "rc=$?"
timeout 200 python tools/diag_step_gaps.py sb_150k_512x288_K9 > gpurun_out/diag_sb.json 2> gpurun_out/diag_sb.err; echo "rc=$?"
