#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the cluster decode kernels, the beam path and the tcgen05 GEMM, on
# short sequences (32 steps).  Logs -> gpurun_out/sanitize_*.log ; summary lines are copied to profiles/ by hand.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {   # tool, tag, args...
  tool=$1; tag=$2; shift 2
  timeout 900 $CS --tool $tool --print-limit 20 python tools/sanitize_case.py "$@" > gpurun_out/sanitize_${tool}_${tag}.log 2>&1
  echo "== $tool $tag: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_${tag}.log | tail -1)"
}
TOOLS=${1:-"memcheck racecheck"}
for tool in $TOOLS; do
  run $tool wide_b5 wide 5
  run $tool cluster16_b5 cluster16 5
  run $tool graph_b5 graph 5
  run $tool beam_b3 beam 3 3
  run $tool gemm gemm 0
  run $tool swin_b2 swin 2
  run $tool convnext_b2 convnext 2
  run $tool tiled_b5 tiled 5
  # multi-cluster cases: racecheck serialises the clusters that spin on row_state (each hit the 900 s limit in round 2)
  if [ $tool = memcheck ]; then
    run $tool wide_b33 wide 33
    run $tool cluster16_b33 cluster16 33
    run $tool cluster_b37 cluster 37
    run $tool auto_b250 auto 250      # throughput kernel selected automatically, two launches (240 + 10 rows)
  fi
done
