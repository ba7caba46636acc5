#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_swin.py tests/test_gpu_shapes.py tests/test_gpu_facade.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/iter_pytest.log
for b in 32 64 128 256; do timeout 300 python tools/quick_enc_bench.py $b 2>&1 | tail -1; done | tee gpurun_out/iter_enc.log
PROFILE_BATCH=256 PROFILE_ENCODE_ONLY=1 timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2g_launches_swin_encoder_b256.csv python tools/profile_step.py > gpurun_out/profile.log 2>&1
python tools/summarize_launches.py gpurun_out/r2g_launches_swin_encoder_b256.csv | tee gpurun_out/r2g_launches_swin_encoder_b256.md
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 2> gpurun_out/bench_c4.err | tee gpurun_out/r2g_bench_c4.json | cut -c1-300
