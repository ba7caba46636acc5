#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_swin.py tests/test_gpu_shapes.py tests/test_gpu_facade.py tests/test_gpu_convnext.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/iter_pytest.log
timeout 300 python tools/quick_enc_bench.py 2>&1 | tail -1 | tee gpurun_out/iter_enc.log
PROFILE_ENCODE_ONLY=1 timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2i_launches_swin_encoder.csv python tools/profile_step.py > gpurun_out/profile.log 2>&1
python tools/summarize_launches.py gpurun_out/r2i_launches_swin_encoder.csv | tee gpurun_out/r2i_launches_swin_encoder.md
timeout 600 python tools/pipe_bench.py 20 7 2>&1 | tail -3 | tee gpurun_out/iter_pipe.log
