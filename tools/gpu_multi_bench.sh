#!/bin/bash
# usage: tools/gpu_round12.sh N  (under gpurun --gpus N): ensemble config (BASELINE config 4) and the default bench at N ranks
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655"
timeout 600 $TR bench.py --gpus $N --workload ensemble --steps 200 --warmup 3 > gpurun_out/r12_bench_ensemble_n$N.json 2> gpurun_out/r12_bench_ensemble_n$N.err; echo "ensemble rc=$?"; tail -1 gpurun_out/r12_bench_ensemble_n$N.json | cut -c1-900
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r12_bench_n$N.json 2> gpurun_out/r12_bench_n$N.err; echo "bench rc=$?"; tail -1 gpurun_out/r12_bench_n$N.json | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r12_pytest_multi_$N.log 2>&1; echo "pytest multi rc=$?"; tail -2 gpurun_out/r12_pytest_multi_$N.log
