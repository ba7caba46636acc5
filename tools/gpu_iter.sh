#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/pipe_bench.py 20 5 6 7 8 2>&1 | tail -7 | tee gpurun_out/iter_pipe.log
timeout 600 python tools/pipe_bench.py 21 7 2>&1 | tail -2 | tee -a gpurun_out/iter_pipe.log
timeout 600 python tools/pipe_bench.py 42 6 7 2>&1 | tail -3 | tee -a gpurun_out/iter_pipe.log
