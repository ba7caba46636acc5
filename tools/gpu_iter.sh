#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_shapes.py tests/test_gpu_swin.py -q -m gpu 2>&1 | tail -6 | tee gpurun_out/iter_pytest.log
MNX_DECODE_PATH=wide timeout 300 python tools/quick_dec_bench.py 32 2>&1 | tail -1 | tee gpurun_out/iter_dec.log
timeout 600 python tools/pipe_bench.py 20 7 2>&1 | tail -3 | tee gpurun_out/iter_pipe.log
