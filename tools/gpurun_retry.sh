#!/bin/bash
# gpurun with retries while the pod has no free GPU slot (nothing is charged for a refused call)
# usage: tools/gpurun_retry.sh <timeout> '<command>'
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "gave up: no GPU slot"; exit 3
