#!/bin/bash
# usage: tools/gpu_multi.sh N [tag]   (run under gpurun --gpus N): multi-GPU tests, then the default bench at N ranks
N=${1:-2}; TAG=${2:-r2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi_$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/${TAG}_pytest_multi_$N.log 2>&1; echo "pytest multi rc=$?"; tail -15 gpurun_out/${TAG}_pytest_multi_$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench_n$N.err; cat gpurun_out/${TAG}_bench_n$N.json
