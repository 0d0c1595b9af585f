#!/bin/bash
# what the driver runs at round end, on one GPU
mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/final_pytest.log
( time timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err ) 2> gpurun_out/final_bench.time; echo "bench rc=$?"; tail -2 gpurun_out/final_bench.err; cat gpurun_out/final_bench.time; cut -c1-600 gpurun_out/final_bench.json
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2>&1 ) 2> gpurun_out/final_ref.time; echo "ref rc=$?"; cut -c1-400 gpurun_out/final_bench_ref.json; cat gpurun_out/final_ref.time
