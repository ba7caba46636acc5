#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_facade.py tests/test_gpu_decoder.py -q -m gpu 2>&1 | tail -8 | tee gpurun_out/iter_pytest.log
