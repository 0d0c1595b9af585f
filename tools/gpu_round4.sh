#!/bin/bash
mkdir -p gpurun_out
{
timeout 200 python tools/probe.py traj b1024 20000
timeout 200 python tools/probe.py traj b512 20000
NB_GRID_MIN_N=16 timeout 200 python tools/probe.py traj b200 20000
NB_GRID_MIN_N=16 timeout 200 python tools/probe.py traj b100 20000
timeout 300 python tools/probe.py solve b1024
timeout 300 python tools/probe.py solve b512
} > gpurun_out/probe4.log 2>&1
cat gpurun_out/probe4.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
