#!/bin/bash
# (GPU box) memcheck of this session's new kernels / paths: depth normals, operand pack, sparse-gradient backward, graphed step
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python -m pytest -q -p no:cacheprovider \
  tests/test_normals_gpu.py "tests/test_hexplane_gpu.py::test_device_operand_pack_is_bit_exact" \
  "tests/test_render_gpu.py::test_render_blurry_view_equals_reference_loop" tests/test_graphs_gpu.py \
  -k "not 1920" > gpurun_out/r2c_san.log 2>&1; echo "rc=$?"
grep -n "Invalid\|by thread\|Address 0x\|=========     at\|ERROR SUMMARY" gpurun_out/r2c_san.log | head -20; tail -4 gpurun_out/r2c_san.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
