#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_swin.py -q -m gpu 2>&1 | tail -3 | tee gpurun_out/iter_pytest.log
timeout 600 python tools/pipe_bench.py 20 2>&1 | tail -3 | tee gpurun_out/iter_pipe.log
