#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_swin.py tests/test_gpu_shapes.py tests/test_gpu_facade.py tests/test_gpu_convnext.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/iter_pytest.log
timeout 300 python tools/quick_enc_bench.py 2>&1 | tail -1 | tee gpurun_out/iter_enc.log
timeout 300 python tools/quick_cn_bench.py 2>&1 | tail -2 | tee gpurun_out/iter_cn.log
# beam step breakdown: skip the precompute + first graph (cold), profile the second graph launch (16 steps x 39 kernels)
timeout 900 ncu --clock-control none --metrics gpu__time_duration.sum -s 700 -c 640 --csv --log-file gpurun_out/r2f_launches_beam1280.csv python tools/profile_beam.py > gpurun_out/profile.log 2>&1
python tools/summarize_launches.py gpurun_out/r2f_launches_beam1280.csv | tee gpurun_out/r2f_launches_beam1280.md
PROFILE_ENCODE_ONLY=1 timeout 600 ncu --clock-control none --set full --import-source on -k regex:window_attn -s 10 -c 1 -f -o gpurun_out/r2f_winattn python tools/profile_step.py >> gpurun_out/profile.log 2>&1
