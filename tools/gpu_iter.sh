#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 7 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/r2r_bench_short.json | cut -c1-160
tail -2 gpurun_out/bench.err
