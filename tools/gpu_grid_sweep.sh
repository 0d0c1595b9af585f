#!/bin/bash
# copy-delay / cluster-size sweep of the grid trajectory kernel (b1024), then the phase profile and the 1-GPU solves
mkdir -p gpurun_out
L=gpurun_out/grid_sweep.log; : > $L
run() { echo "== $*" >> $L; timeout 200 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
for cs in 4 2; do
for d in 600 800 900 1000 1200; do
echo "### one system per launch, clusters of $cs, copy delay $d clk" >> $L
NB_GRID_CS=$cs NB_GRID_DELAY=$d run python tools/probe.py trajrep b1024 100000 2
done
done
echo "### phase profile, defaults" >> $L
NB_GRID_PROFILE=1 run python tools/probe.py traj b1024 20000
for d in 900 2000 2600 3200; do
echo "### three-query solve on one GPU (two systems per launch), copy delay $d clk" >> $L
NB_GRID_DELAY=$d run python tools/probe.py solve b1024
done
run python tools/probe.py solve b512
run python tools/probe.py solve b200
cat $L
