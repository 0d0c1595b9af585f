#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r9.log; : > $L
run() { echo "== $*" >> $L; timeout 300 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
export NB_VERBOSE=1
run python tools/probe.py solve b1024
run python tools/probe.py solve b1024
run python tools/probe.py solve b512
cat $L
