#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | tee gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_parallel.py -q -m gpu -s 2>&1 | tail -6 | tee gpurun_out/r2i_pytest_parallel.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/bench2.err | tee gpurun_out/r2i_bench_2gpu.json | cut -c1-300
tail -2 gpurun_out/bench2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c4 --steps 5 --warmup 3 2> gpurun_out/bench2c4.err | tee gpurun_out/r2i_bench_c4_2gpu.json | cut -c1-300
tail -2 gpurun_out/bench2c4.err
