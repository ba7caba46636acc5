#!/bin/bash
# One GPU-box visit: tests, smoke, bench (both arms), ncu launch list. Outputs under gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json
python bench.py --impl reference --steps 2 --warmup 3 2>> gpurun_out/bench.err | tee gpurun_out/bench_ref.json
tail -5 gpurun_out/bench.err
