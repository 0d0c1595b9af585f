#!/bin/bash
L=gpurun_out/s3_e.log; : > $L
for v in base swp2 swp2s sleep swp3 swp4; do
  for k in "NB_GRID_SPLIT=0" "NB_GRID_SPLIT=1"; do
    echo "== $v $k" >> $L
    env $k NB_LIB_PATH=tools/_build/variants/lib_$v.so python tools/grid_profile.py b1024 2>&1 | grep -v "^grid" >> $L
  done
done
cat $L
