#!/bin/bash
# round-2 verification pass: whole GPU suite, smoke, bench (metric config c2, both arms) and the other BASELINE configs
mkdir -p gpurun_out
tag=${1:-r2f}
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | tee gpurun_out/gpu.txt
timeout 2400 python -m pytest tests -q -m gpu -x -s 2>&1 | grep -vE "^\s*$" | tail -40 | tee gpurun_out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.log
timeout 600 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/${tag}_bench.json | cut -c1-300
tail -3 gpurun_out/bench.err
for c in c1 c3 c4 c5; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 2> gpurun_out/bench_$c.err | tee gpurun_out/${tag}_bench_$c.json | cut -c1-400
  tail -2 gpurun_out/bench_$c.err
done
timeout 900 python bench.py --impl reference 2> gpurun_out/bench_ref.err | tee gpurun_out/${tag}_bench_ref.json | cut -c1-300
