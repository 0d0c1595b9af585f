#!/bin/bash
# grid kernel A/B on one box: round-1 build, the full-record validation build, the current build
L=${1:-gpurun_out/grid_ab.log}; : > $L
for lib in tools/_build/keep/lib_r1.so tools/_build/keep/lib_grid_unpiped.so nthu_ipc_nbody-simulation_b200/libnbody_b200.so; do
  echo "== $lib" >> $L
  for d in 900 1200; do
    echo "-- delay $d" >> $L
    NB_LIB_TOLERANT=1 NB_LIB_PATH=$PWD/$lib NB_GRID_DELAY=$d python tools/grid_profile.py b1024 >> $L 2>&1
  done
  NB_LIB_TOLERANT=1 NB_LIB_PATH=$PWD/$lib NB_GRID_PROFILE=1 python tools/grid_profile.py b1024 40000 2>&1 | grep "grid profile" >> $L
  NB_LIB_TOLERANT=1 NB_LIB_PATH=$PWD/$lib python tools/grid_profile.py b512 >> $L 2>&1
done
python -c "
import importlib,sys; sys.path.insert(0,'.'); nb=importlib.import_module('nthu_ipc_nbody-simulation_b200')
s=nb.read_input('tests/golden/testcases/b1024.in'); nb.solve(s,gpus=[0]); print('torn records:', nb.grid_torn_records())" >> $L 2>&1
cat $L
