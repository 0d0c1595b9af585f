#!/bin/bash
# hw5 process wall time: default GPU choice vs one GPU, on a multi-GPU box
mkdir -p gpurun_out
L=gpurun_out/r11.log; : > $L
T=tests/golden/testcases
H=nthu_ipc_nbody-simulation_b200/hw5
t() { local s=$(date +%s.%N); "$@" >> $L 2>&1; local rc=$?; local e=$(date +%s.%N); echo "wall $(echo "$e - $s" | bc) s rc=$rc :: $*" >> $L; }
nvidia-smi -L >> $L
for rep in 1 2; do
NB_VERBOSE=1 t $H $T/b1024.in /tmp/o1.out; cmp /tmp/o1.out $T/b1024.out >> $L && echo identical >> $L
CUDA_VISIBLE_DEVICES=0 NB_VERBOSE=1 t $H $T/b1024.in /tmp/o2.out; cmp /tmp/o2.out $T/b1024.out >> $L && echo identical >> $L
done
t $H $T/b200.in /tmp/o3.out; cmp /tmp/o3.out $T/b200.out >> $L && echo identical >> $L
t $H $T/b20.in /tmp/o4.out; cmp /tmp/o4.out $T/b20.out >> $L && echo identical >> $L
grep -v "Q2 chunk\|host spent" $L
