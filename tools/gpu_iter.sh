#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/r2p_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r2p_smoke.log
timeout 300 python tools/quick_dec_bench.py 32 2>&1 | tail -1 | tee gpurun_out/iter_dec.log
