#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/grid3.log; : > $L
run() { echo "== $*" >> $L; timeout 120 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
for cs in 4 2; do
for d in 600 800 900 1000 1200; do
echo "### v2 CS=$cs delay=$d (no profile, 100k steps)" >> $L
NB_GRID_CS=$cs NB_GRID_DELAY=$d run python tools/probe.py traj b1024 100000
done
done
echo "### profile, defaults" >> $L
NB_GRID_PROFILE=1 run python tools/probe.py traj b1024 20000
echo "### v1" >> $L
NB_GRID_IMPL=1 run python tools/probe.py traj b1024 100000
echo "### solves, defaults" >> $L
run python tools/probe.py solve b200
run python tools/probe.py solve b512
run python tools/probe.py solve b1024
NB_GRID_IMPL=1 run python tools/probe.py solve b1024
grep -v "^Traceback\|^  File\|^    " $L
