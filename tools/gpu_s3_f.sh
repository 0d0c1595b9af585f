#!/bin/bash
L=gpurun_out/s3_f.log; : > $L
for k in "NB_GRID_CS=4" "NB_GRID_CS=8" "NB_GRID_CS=2" "NB_GRID_CS=8 NB_GRID_ADAPT=0,0 NB_GRID_DELAY=700" "NB_GRID_CS=8 NB_GRID_ADAPT=0,0 NB_GRID_DELAY=1000"; do
  echo "== $k" >> $L
  env $k python tools/grid_profile.py b1024 100000 2>&1 | grep "Q1" >> $L
  env $k NB_GRID_PROFILE=1 NB_GRID_T=1 python tools/grid_profile.py b1024 20000 2>&1 | grep "grid profile" | sed -n '2p' | cut -c1-300 >> $L
done
cat $L
