#!/bin/bash
# memcheck of the kernels added after the first sanitizer pass
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() { tool=$1; tag=$2; shift 2
  timeout 300 $CS --tool $tool --print-limit 20 python tools/sanitize_case.py "$@" > gpurun_out/sanitize_${tool}_${tag}.log 2>&1
  echo "== $tool $tag: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_${tag}.log | tail -1) | $(tail -1 gpurun_out/sanitize_${tool}_${tag}.log | cut -c1-100)"
}
run memcheck swin_b2 swin 2
run memcheck convnext_b2 convnext 2
run memcheck tiled_b5 tiled 5
run memcheck auto_b250 auto 250
run racecheck tiled_b5 tiled 5
run racecheck swin_b2 swin 2
