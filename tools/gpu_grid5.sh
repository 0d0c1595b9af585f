#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/grid5.log; : > $L
run() { echo "== $*" >> $L; timeout 300 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
NB_GRID_PROFILE=1 run python tools/probe.py trajrep b1024 50000 3
run python tools/probe.py trajrep b1024 100000 4
run python tools/probe.py solve b1024
run python tools/probe.py solve b512
run python tools/probe.py solve b200
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "grid or solve or trajectory or nbtool" >> $L 2>&1
cat $L
