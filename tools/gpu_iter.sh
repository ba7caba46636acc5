#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_decoder.py tests/test_gpu_shapes.py -q -m gpu -s 2>&1 | grep -vE "^\s*$" | tail -12 | tee gpurun_out/iter_pytest.log
for b in 128 256; do echo "== auto B=$b"; timeout 300 python tools/quick_dec_bench.py $b 2>&1 | tail -1; done | tee gpurun_out/iter_dec_scale.log
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 2> gpurun_out/bench_c4.err | tee gpurun_out/r2j_bench_c4.json | cut -c1-250
tail -2 gpurun_out/bench_c4.err
