#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c4 --steps 5 --warmup 3 2> gpurun_out/bench2c4.err | tee gpurun_out/r2q_bench_c4_2gpu.json | cut -c1-250
tail -2 gpurun_out/bench2c4.err
