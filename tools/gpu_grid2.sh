#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/grid2.log; : > $L
run() { echo "== $*" >> $L; timeout 90 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
for ns in 0 50 200 500; do
export NB_GRID_POLL_NS=$ns
echo "### poll sleep $ns ns" >> $L
run python tools/probe.py traj b1024 200000
run python tools/probe.py solve b1024
done
cat $L
