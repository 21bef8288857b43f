#!/bin/bash
# Round-2 ablation on the headline workload (runs on the GPU box): one decision of this round switched off at a time.
# Library variants are prebuilt on the CPU box into build/variants/lib_<name>.so:
#   product (none)   noleanA -DMOBGS_BWD_LEAN_A=0   nomom -DMOBGS_BWD_UNIT_MOMENTS=0   ranksort -DMOBGS_RANK_SORT_BUCKETS=0
mkdir -p gpurun_out
run() {  # $1 = label, $2 = variant, $3 = env, $4 = extra bench flags
  echo "=== $1"
  cp build/variants/lib_$2.so mobgs_b200/libmobgs_b200.so
  env $3 python bench.py --no-cpu-baseline --no-extras --steps 15 --warmup 5 $4 2>/dev/null | python tools/show_bench.py | sed -n 1,2p
}
cp mobgs_b200/libmobgs_b200.so /tmp/lib_keep.so
run "product build" product X=1
run "backward Phase A with per-iteration list-end / blended tests (no dummy-record padding)" noleanA X=1
run "backward Phase B: moments accumulated about the mean pixel by pixel" nomom X=1
run "all-pairs rank sort instead of the bucketed rank sort" ranksort X=1
run "two-pass count / emit kernels (no recorded entries)" product MOBGS_RECORD_ENTRIES=0
run "gradient-record buffer allocated + zero-filled every step" product MOBGS_RECYCLE_GRAD_RECORDS=0
run "ray images [K,6,H,W] (camera_rays kernels) instead of in-register rays" product X=1 --ray-images
cp /tmp/lib_keep.so mobgs_b200/libmobgs_b200.so
