#!/bin/bash
# Ablation of the blend-path design decisions on the headline workload (runs on the GPU box):
# each line rebuilds the library with one optimisation switched off and prints the bench kernel table.
set -e
run() {  # $1 = label, $2 = nvcc flags, $3 = env
  echo "=== $1"
  MOBGS_NVCC_EXTRA="$2" python -c "from mobgs_b200 import _lib; _lib.build(force=True)" 2>/dev/null
  env $3 python bench.py --no-cpu-baseline --steps 15 2>/dev/null | python tools/show_bench.py | head -3
}
run "product build" "" "X=1"
run "no strip masks" "-DMOBGS_ABL_NO_STRIP_MASK" "X=1"
run "naive 5-step shuffle reduction" "-DMOBGS_ABL_NAIVE_REDUCE" "X=1"
run "LDG/STS staging instead of TMA bulk copies" "-DMOBGS_TMA_STAGE=0" "X=1"
run "gsplat tile AABB (no exact tile pruning)" "" "MOBGS_ABL_AABB_TILES=1"
run "all of the above off" "-DMOBGS_ABL_NO_STRIP_MASK -DMOBGS_ABL_NAIVE_REDUCE -DMOBGS_TMA_STAGE=0" "MOBGS_ABL_AABB_TILES=1"
MOBGS_NVCC_EXTRA="" python -c "from mobgs_b200 import _lib; _lib.build(force=True)" 2>/dev/null
