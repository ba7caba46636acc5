#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_swin.py tests/test_gpu_convnext.py tests/test_gpu_facade.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/iter_pytest.log
timeout 300 python tools/quick_enc_bench.py 2>&1 | tail -1 | tee gpurun_out/iter_enc.log
timeout 600 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/r2l_bench.json | cut -c1-250
timeout 600 python bench.py --config c1 --steps 5 --warmup 3 2> gpurun_out/bench_c1.err | tee gpurun_out/r2l_bench_c1.json | cut -c1-250
