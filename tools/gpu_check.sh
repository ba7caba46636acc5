#!/bin/bash
# one GPU round: beam parity first (newest code), then the whole GPU suite, smoke, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | tee gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_gpu_beam.py -x -q -s -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_beam.log
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_beam.py 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-400
tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 2>> gpurun_out/bench.err | tee gpurun_out/bench_ref.json | cut -c1-300
