#!/bin/bash
# full single-GPU pass with the new grid kernel: smoke, gpu tests, bench (both arms)
mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r6_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r6_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r6_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r6_pytest.log
timeout 600 python bench.py > gpurun_out/r6_bench.json 2> gpurun_out/r6_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r6_bench.err; cat gpurun_out/r6_bench.json
