#!/bin/bash
# A/B of grid-kernel builds (tools/build_grid_variants.sh): one-system step latency and the one-GPU three-query solve
L=${L:-gpurun_out/grid_variants.log}; : > $L
for v in "$@"; do
  echo "== $v" >> $L
  NB_LIB_PATH=tools/_build/variants/lib_$v.so timeout 120 python tools/grid_profile.py b1024 2>&1 | grep -v "^grid" >> $L
  NB_LIB_PATH=tools/_build/variants/lib_$v.so NB_GRID_PROFILE=1 timeout 120 python tools/grid_profile.py b1024 20000 2>&1 | grep "grid profile" | sed -n '2p;5p' | cut -c1-300 >> $L
done
cat $L
