#!/bin/bash
# new tests + ncu capture of the grid kernel (exchange v2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r8_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r8_pytest.log
GRD="python tools/probe.py traj b1024 2000"
timeout 200 $GRD > gpurun_out/r8_plain_grid.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grid_traj -s 1 -c 1 -o gpurun_out/r8_prof_grid_kernel -f $GRD > gpurun_out/r8_ncu_grid.log 2>&1
echo "ncu grid rc=$?"; cat gpurun_out/r8_plain_grid.log; tail -3 gpurun_out/r8_ncu_grid.log
