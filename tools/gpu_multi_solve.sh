#!/bin/bash
# usage: tools/gpu_round7.sh N   (run under gpurun --gpus N): multi-GPU tests, bench at N, hw5 CLI wall time
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r7_smi_$N.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r7_pytest_multi_$N.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/r7_pytest_multi_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r7_bench_n$N.json 2> gpurun_out/r7_bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/r7_bench_n$N.err; cat gpurun_out/r7_bench_n$N.json
(timeout 200 python tools/probe.py solve b1024 $N; timeout 200 python tools/probe.py solve b1024 $N; timeout 200 python tools/probe.py cli b1024 2; timeout 100 python tools/probe.py solve b512 $N) > gpurun_out/r7_cli_$N.log 2>&1; cat gpurun_out/r7_cli_$N.log
