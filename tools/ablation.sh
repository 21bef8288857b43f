#!/bin/bash
# Ablation of the blend-path design decisions on the headline workload (runs on the GPU box).  The library
# variants are prebuilt on the CPU box into build/variants/lib_<name>.so (flags in the table below):
#   product  (none)                         nomask   -DMOBGS_ABL_NO_STRIP_MASK      (every unit visits every entry)
#   bfly     -DMOBGS_BWD_TRANSPOSE=0        naive    -DMOBGS_BWD_TRANSPOSE=0 -DMOBGS_ABL_NAIVE_REDUCE
#   smemacc  -DMOBGS_BWD_DIRECT_RED=0       L32 / L8 -DMOBGS_UNIT_LANES=32 / 8     (8x4 / 4x2-pixel units)
#   notma    -DMOBGS_TMA_STAGE=0            (LDG/STS staging in the forward; the transposing backward always uses TMA)
mkdir -p gpurun_out
run() {  # $1 = label, $2 = variant, $3 = env
  echo "=== $1"
  cp build/variants/lib_$2.so mobgs_b200/libmobgs_b200.so
  env $3 python bench.py --no-cpu-baseline --steps 15 2>/dev/null | python tools/show_bench.py | head -2
}
run "product build" product X=1
run "no unit masks (every unit walks every list entry)" nomask X=1
run "backward: butterfly reduction instead of ownership transposition" bfly X=1
run "backward: naive log-step shuffle reduction per value" naive X=1
run "backward: shared-memory accumulator + flush instead of direct vector reductions" smemacc X=1
run "units of 32 lanes (8x4 px, butterfly backward)" L32 X=1
run "units of 8 lanes (4x2 px, butterfly backward)" L8 X=1
run "forward: LDG/STS staging instead of TMA bulk copies" notma X=1
run "gsplat tile AABB (no exact tile pruning)" product MOBGS_ABL_AABB_TILES=1
cp build/variants/lib_product.so mobgs_b200/libmobgs_b200.so
