#!/bin/bash
# usage: tools/ab_build.sh "<nvcc -D flags variant 1>" "<variant 2>" ...   (runs on the GPU box)
# rebuilds the library with each flag set and prints the bench kernel table.
for v in "$@"; do
  echo "=== variant: $v"
  MOBGS_NVCC_EXTRA="$v" python -c "from mobgs_b200 import _lib; _lib.build(force=True)" || exit 1
  python bench.py --no-cpu-baseline --steps 20 2>&1 | python tools/show_bench.py | head -3
done
