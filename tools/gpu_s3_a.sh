#!/bin/bash
# session-3 probe A: does the adaptive copy delay hurt the two-systems-per-launch solve?
L=gpurun_out/s3_a.log; : > $L
for a in "128,8" "0,0"; do
  echo "== NB_GRID_ADAPT=$a" >> $L
  NB_GRID_ADAPT=$a python tools/grid_profile.py b1024 2>&1 | grep -v "^grid" >> $L
  NB_GRID_ADAPT=$a NB_GRID_PROFILE=1 python tools/grid_profile.py b1024 20000 2>&1 | grep "grid profile" | sed -n '2p;5p' >> $L
done
cat $L
