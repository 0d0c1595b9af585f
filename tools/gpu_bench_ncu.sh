#!/bin/bash
# bench line + ncu launch list + one full ncu capture of the dominant kernel
mkdir -p gpurun_out
SHORT="python bench.py --steps 2 --warmup 1 --no-b1024 --no-cpu-baseline"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/bench_n1.json
timeout 300 $SHORT > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 300 $SHORT > gpurun_out/plain2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:large_accel -s 2 -c 2 -o gpurun_out/prof_large_accel -f $SHORT > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full.log
