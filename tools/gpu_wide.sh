#!/bin/bash
# wide (throughput) decode kernel: parity, then per-step timing alone
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_shapes.py -x -q -m gpu -k "wide" 2>&1 | tail -15
for b in "$@"; do MNX_DECODE_PATH=wide MNX_DECODE_PROFILE=1 timeout 120 python tools/quick_dec_bench.py $b 2>&1 | tail -30; done
if [ -n "$PIPE" ]; then timeout 600 python tools/pipe_bench.py $PIPE 2>&1 | tail -12; fi
