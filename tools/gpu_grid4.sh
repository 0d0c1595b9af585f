#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/grid4.log; : > $L
run() { echo "== $*" >> $L; timeout 200 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
for d in 900 1400 2000 2600 3200; do
echo "### two systems per launch, copy delay $d" >> $L
NB_GRID_DELAY=$d run python tools/probe.py solve b1024
NB_GRID_DELAY=$d run python tools/probe.py solve b512
done
cat $L
