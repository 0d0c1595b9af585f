#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
L=gpurun_out/r5_$N.log; : > $L
run() { echo "== $*" >> $L; timeout 120 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python tools/probe.py cli b1024 2
run python tools/probe.py cli b200 1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 20 --warmup 3 --no-b1024 > gpurun_out/bench_n${N}b.json 2> gpurun_out/bench_n${N}b.err; echo "bench rc=$?" >> $L
cat $L; cut -c1-900 gpurun_out/bench_n${N}b.json
