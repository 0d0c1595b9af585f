#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
echo skip-pytest
for ex in nccl p2p; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 30 --warmup 3 --no-b1024 --exchange $ex > gpurun_out/bench_n${N}_$ex.json 2> gpurun_out/bench_n${N}_$ex.err; echo "bench $ex rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n${N}_$ex.json") if l.startswith("{")][-1])
print("$ex", "value %.4e" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"], "frac %.3f" % d["frac_of_fp64_peak"])
PY
done
