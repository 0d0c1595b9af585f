#!/bin/bash
mkdir -p gpurun_out
ENS="python bench.py --workload ensemble --systems 148 --steps 50 --warmup 1"
GRD="python tools/probe.py traj b1024 2000"
timeout 300 python bench.py --workload ensemble --steps 200 --warmup 2 > gpurun_out/bench_ensemble_n1.json 2> gpurun_out/bench_ensemble_n1.err; echo "ensemble bench rc=$?"; cat gpurun_out/bench_ensemble_n1.json
timeout 200 $ENS > gpurun_out/plain_ens.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:traj_kernel -s 1 -c 1 -o gpurun_out/prof_traj_kernel -f $ENS > gpurun_out/ncu_ens.log 2>&1
echo "ncu ens rc=$?"
timeout 200 $GRD > gpurun_out/plain_grid.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grid_traj -s 1 -c 1 -o gpurun_out/prof_grid_kernel -f $GRD > gpurun_out/ncu_grid.log 2>&1
echo "ncu grid rc=$?"
tail -2 gpurun_out/ncu_ens.log gpurun_out/ncu_grid.log
