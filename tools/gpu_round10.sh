#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r10.log; : > $L
run() { echo "== $*" >> $L; timeout 300 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
for d in 600 700 800 900 1000 1100; do
echo "### T=1 delay $d" >> $L
NB_GRID_DELAY=$d run python tools/probe.py trajrep b1024 100000 2
done
for d in 900 1500 2000 2600 3200 4000; do
echo "### solve (chain, two systems per launch) delay $d" >> $L
NB_GRID_DELAY=$d run python tools/probe.py solve b1024
done
for c in 2048 4096 16384; do
echo "### chunk $c" >> $L
NB_SOLVE_CHUNK=$c run python tools/probe.py solve b1024
done
cat $L
