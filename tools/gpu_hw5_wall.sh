#!/bin/bash
# hw5 process wall time against the floor of any CUDA process on this box
mkdir -p gpurun_out
L=${1:-gpurun_out/hw5_wall.log}; : > $L
T=tests/golden/testcases
P=nthu_ipc_nbody-simulation_b200
t() { local s=$(date +%s.%N); "$@" >> $L 2>&1; local rc=$?; local e=$(date +%s.%N); echo "wall $(python3 -c "print(round($e - $s, 3))") s rc=$rc :: $*" >> $L; }
nvidia-smi -L >> $L
for rep in 1 2 3; do
t $P/cuda_floor
NB_VERBOSE=1 t $P/hw5 $T/b1024.in /tmp/o1.out; cmp /tmp/o1.out $T/b1024.out >> $L && echo identical >> $L
done
t $P/hw5 $T/b200.in /tmp/o3.out; cmp /tmp/o3.out $T/b200.out >> $L && echo identical >> $L
t $P/hw5 $T/b20.in /tmp/o4.out; cmp /tmp/o4.out $T/b20.out >> $L && echo identical >> $L
grep -v "Q2 chunk\|host spent" $L
