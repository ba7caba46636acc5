#!/bin/bash
# quick decode iteration: parity of the decode paths, then per-step timing at the given batch sizes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_shapes.py -x -q -m gpu 2>&1 | tail -4
for b in "$@"; do MNX_DECODE_PROFILE=1 timeout 120 python tools/quick_dec_bench.py $b 2>&1 | grep -v "^max co" | tail -30; done
