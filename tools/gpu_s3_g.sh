#!/bin/bash
L=gpurun_out/s3_g.log; : > $L
for v in "$@"; do
    echo "== $v" >> $L
    NB_LIB_PATH=tools/_build/variants/lib_$v.so python tools/grid_profile.py b1024 2>&1 | grep -v "^grid" >> $L
    NB_LIB_PATH=tools/_build/variants/lib_$v.so NB_GRID_PROFILE=1 python tools/grid_profile.py b1024 20000 2>&1 | grep "grid profile" | sed -n '2p;5p' | cut -c1-300 >> $L
done
cat $L
