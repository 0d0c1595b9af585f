#!/bin/bash
# usage: tools/gpu_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_smi_$N.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_$N.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_multi_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
