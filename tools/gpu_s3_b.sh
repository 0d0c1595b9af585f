#!/bin/bash
# session-3 probe B: constant copy delay sweep for the two-systems-per-launch solve (adaptive off)
L=gpurun_out/s3_b.log; : > $L
for d in 2600 3400 4200 5000 6000 7000; do
  echo "== NB_GRID_ADAPT=0,0 NB_GRID_DELAY=$d" >> $L
  NB_GRID_ADAPT=0,0 NB_GRID_DELAY=$d python tools/grid_profile.py b1024 40000 2>&1 | grep "three-query" >> $L
  NB_GRID_ADAPT=0,0 NB_GRID_DELAY=$d NB_GRID_PROFILE=1 python tools/grid_profile.py b1024 20000 2>&1 | grep "grid profile" | sed -n '5p' >> $L
done
cat $L
