#!/bin/bash
L=${1:-gpurun_out/grid_ab2.log}; : > $L
for lib in tools/_build/keep/lib_grid_prev.so nthu_ipc_nbody-simulation_b200/libnbody_b200.so tools/_build/keep/lib_grid_u1.so tools/_build/keep/lib_grid_u4.so; do
  echo "== $lib" >> $L
  NB_LIB_TOLERANT=1 NB_LIB_PATH=$PWD/$lib NB_GRID_PROFILE=1 python tools/grid_profile.py b1024 40000 2>&1 | grep "grid profile T=1" | tail -1 >> $L
  NB_LIB_TOLERANT=1 NB_LIB_PATH=$PWD/$lib python tools/grid_profile.py b1024 >> $L 2>&1
  NB_LIB_TOLERANT=1 NB_LIB_PATH=$PWD/$lib python tools/grid_profile.py b512 2>&1 | head -1 >> $L
  NB_LIB_TOLERANT=1 NB_LIB_PATH=$PWD/$lib python tools/grid_profile.py b200 2>&1 | head -1 >> $L
done
cat $L
