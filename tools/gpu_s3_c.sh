#!/bin/bash
# session-3 probe C: pair-phase time with 4 compute warps (NJ=2) against 8 (NJ=4), one system
L=gpurun_out/s3_c.log; : > $L
for nj in 4 2; do
  echo "== NB_GRID_NJ=$nj" >> $L
  NB_GRID_NJ=$nj python tools/grid_profile.py b1024 40000 2>&1 | grep "Q1" >> $L
  NB_GRID_NJ=$nj NB_GRID_PROFILE=1 NB_GRID_T=1 python tools/grid_profile.py b1024 20000 2>&1 | grep "grid profile" | sed -n '2p' >> $L
done
cat $L
