#!/bin/bash
# round-2 verification pass: whole GPU suite, smoke, bench (both arms), encoder launch list under ncu
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | tee gpurun_out/gpu.txt
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 | tee gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/r2b_bench.json | cut -c1-300
tail -3 gpurun_out/bench.err
PROFILE_ENCODE_ONLY=1 timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2b_launches_swin_encoder.csv python tools/profile_step.py > gpurun_out/profile.log 2>&1
python tools/summarize_launches.py gpurun_out/r2b_launches_swin_encoder.csv | tee gpurun_out/r2b_launches_swin_encoder.md
