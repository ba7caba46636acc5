#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -3 gpurun_out/bench.err
PROFILE_ENCODER=convnext_base PROFILE_ENCODE_ONLY=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_convnext.csv python tools/profile_step.py > gpurun_out/profile_cn.log 2>&1
PROFILE_ENCODER=convnext_base PROFILE_ENCODE_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv_ln -s 8 -c 1 -f -o gpurun_out/r1_dwconv python tools/profile_step.py >> gpurun_out/profile_cn.log 2>&1
tail -2 gpurun_out/profile_cn.log
