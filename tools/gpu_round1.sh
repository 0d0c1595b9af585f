#!/bin/bash
# first GPU pass: smoke, a few probes, then the gpu tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
(timeout 300 python __graft_entry__.py smoke) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
{
timeout 120 python tools/probe.py peak
timeout 200 python tools/probe.py traj b1024 2000
timeout 200 python tools/probe.py traj b200 20000
timeout 200 python tools/probe.py ens 148 300
for ipt in 1 2 4; do for js in 1 4 8 16; do NB_LARGE_IPT=$ipt NB_LARGE_JSPLIT=$js timeout 200 python tools/probe.py large 65536 10; done; done
timeout 100 python tools/probe.py solve b20
timeout 100 python tools/probe.py solve b200
} > gpurun_out/probe.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/smoke.log; cat gpurun_out/probe.log; tail -15 gpurun_out/pytest_gpu.log
