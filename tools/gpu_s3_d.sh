#!/bin/bash
# session-3 probe D: split warp sets (two systems per launch, each at its own pace) against the lock-step kernel
L=gpurun_out/s3_d.log; : > $L
for k in "NB_GRID_SPLIT=1" "NB_GRID_SPLIT=0" "NB_GRID_SPLIT=0 NB_GRID_ADAPT=0,0 NB_GRID_DELAY=3400"; do
  echo "== $k" >> $L
  env $k python tools/grid_profile.py b1024 2>&1 | grep -v "^grid" >> $L
  env $k NB_GRID_PROFILE=1 python tools/grid_profile.py b1024 20000 2>&1 | grep "grid profile" | sed -n '2p;5p' >> $L
done
env NB_GRID_SPLIT=1 python tools/grid_profile.py b512 2>&1 | grep -v "^grid" >> $L
env NB_GRID_SPLIT=0 python tools/grid_profile.py b512 2>&1 | grep -v "^grid" >> $L
cat $L
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "chain or scheduler or grid or golden" 2>&1 | tail -5
