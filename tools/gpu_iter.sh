#!/bin/bash
# last check of the round: whole GPU suite, smoke, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/r2o_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r2o_smoke.log
timeout 600 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/r2o_bench.json | cut -c1-200
tail -2 gpurun_out/bench.err
