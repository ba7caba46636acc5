#!/bin/bash
# One GPU-box visit: tests, smoke, bench (both arms), ncu launch list + full captures. Outputs under gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json
python bench.py --impl reference --steps 2 --warmup 3 2>> gpurun_out/bench.err | tee gpurun_out/bench_ref.json
tail -5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_cluster.csv python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 60 -c 2 -f -o gpurun_out/r1_gemm_tc python tools/profile_step.py >> gpurun_out/profile_step.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_mega -c 1 -f -o gpurun_out/r1_mega python tools/profile_step.py >> gpurun_out/profile_step.log 2>&1
  MNX_DECODE_PATH=graph timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 41 -c 2 -f -o gpurun_out/r1_xattn python tools/profile_step.py >> gpurun_out/profile_step.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_attn -s 10 -c 1 -f -o gpurun_out/r1_winattn python tools/profile_step.py >> gpurun_out/profile_step.log 2>&1
  tail -3 gpurun_out/profile_step.log; ls -la gpurun_out/
fi
