#!/bin/bash
# scratch iteration script: parity of the changed kernels first, then profiles / timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_convnext.py tests/test_gpu_swin.py tests/test_gpu_facade.py -q -m gpu -s 2>&1 | grep -vE "^\s*$" | tail -30 | tee gpurun_out/iter_pytest.log
NCU="ncu --clock-control none --set full --import-source on"
PROFILE_ENCODER=convnext_base PROFILE_ENCODE_ONLY=1 timeout 600 $NCU -k regex:dwconv_stats -s 8 -c 1 -f -o gpurun_out/r2b_dwconv python tools/profile_step.py > gpurun_out/profile.log 2>&1
PROFILE_ENCODE_ONLY=1 timeout 600 $NCU -k regex:window_attn -s 10 -c 1 -f -o gpurun_out/r2b_winattn python tools/profile_step.py >> gpurun_out/profile.log 2>&1
PROFILE_ENCODE_ONLY=1 timeout 600 $NCU -k regex:gemm_tc_kernel -s 60 -c 3 -f -o gpurun_out/r2b_gemm_tc python tools/profile_step.py >> gpurun_out/profile.log 2>&1
tail -2 gpurun_out/profile.log
timeout 600 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/r2c_bench.json | cut -c1-300
tail -3 gpurun_out/bench.err
