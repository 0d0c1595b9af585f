#!/bin/bash
L=${1:-gpurun_out/grid_adapt.log}; : > $L
for a in "0,0" "64,8" "128,8" "32,4" "200,2" "64,32"; do
  for d in 900 400; do
    echo "== adapt $a start delay $d" >> $L
    NB_GRID_ADAPT=$a NB_GRID_DELAY=$d python tools/grid_profile.py b1024 2>&1 >> $L
  done
done
NB_GRID_ADAPT=64,8 NB_GRID_PROFILE=1 python tools/grid_profile.py b1024 40000 2>&1 | grep "grid profile T=1" | tail -1 >> $L
NB_GRID_ADAPT=64,8 python tools/grid_profile.py b512 >> $L 2>&1
NB_GRID_ADAPT=64,8 python tools/grid_profile.py b200 >> $L 2>&1
cat $L
