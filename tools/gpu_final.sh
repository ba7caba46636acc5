#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 3 2>> gpurun_out/bench.err | tee gpurun_out/bench_ref.json | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_cluster.csv python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 60 -c 2 -f -o gpurun_out/r1_gemm_tc python tools/profile_step.py >> gpurun_out/profile_step.log 2>&1
tail -2 gpurun_out/profile_step.log
