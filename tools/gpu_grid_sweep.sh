#!/bin/bash
# grid trajectory kernel (b1024): pipeline depth x copy delay sweep, phase profiles, 1-GPU solves
L=${1:-gpurun_out/grid_sweep.log}; : > $L
echo "== previous build (un-pipelined, self-checking tags)" >> $L
NB_LIB_PATH=$PWD/tools/_build/keep/lib_grid_unpiped.so python tools/grid_profile.py b1024 >> $L 2>&1
for g in 1 2 4; do for d in 600 900 1200; do
  echo "== groups $g delay $d" >> $L
  NB_GRID_GROUPS=$g NB_GRID_DELAY=$d python tools/grid_profile.py b1024 100000 2>&1 | head -1 >> $L
done; done
for g in 1 2 4; do
  echo "== phase profile groups $g" >> $L
  NB_GRID_GROUPS=$g NB_GRID_PROFILE=1 python tools/grid_profile.py b1024 40000 2>&1 | grep "grid profile T=1" | tail -1 >> $L
done
for g in 1 4; do
  echo "== full runs, groups $g" >> $L
  NB_GRID_GROUPS=$g python tools/grid_profile.py b1024 >> $L 2>&1
  NB_GRID_GROUPS=$g python tools/grid_profile.py b512 >> $L 2>&1
  NB_GRID_GROUPS=$g python tools/grid_profile.py b200 >> $L 2>&1
done
cat $L
