#!/bin/bash
# wide decode kernel: parity (fast subset), timing with phase marks, then one ncu --set full capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decoder.py -x -q -m gpu -k "wide" 2>&1 | tail -3
MNX_DECODE_PATH=wide MNX_DECODE_PROFILE=1 timeout 120 python tools/quick_dec_bench.py 32 2>&1 | tail -30
MNX_DECODE_PATH=wide timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_wide -c 1 -f -o gpurun_out/${1:-r2_wide} python tools/quick_dec_bench.py 32 2>&1 | tail -5
