#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/grid5.log; : > $L
run() { echo "== $*" >> $L; timeout 300 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
NB_GRID_PROFILE=1 run python tools/probe.py trajrep b1024 50000 10
cat $L
