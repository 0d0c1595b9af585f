#!/bin/bash
# grid kernel check: short, every command under its own timeout, logs kept
mkdir -p gpurun_out
L=gpurun_out/grid.log; : > $L
run() { echo "== $*" >> $L; timeout 90 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
NB_GRID_PROFILE=1 run python tools/probe.py traj b1024 200000
run python tools/probe.py traj b1024 200000
run python tools/probe.py traj b512 100000
NB_GRID_MIN_N=16 run python tools/probe.py traj b100 100000
run python tools/probe.py solve b1024
run python tools/probe.py solve b512
cat $L
if grep -q "rc=1\|rc=124" $L; then echo "SKIPPING pytest (a probe failed)"; else
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log; fi
