#!/bin/bash
mkdir -p gpurun_out
(
echo "== default (tile 8,8,8,12 split 1,1,1,8)"; timeout 300 python tools/quick_cn_bench.py 2>&1 | tail -2
for sp in "1,2,2,8" "2,4,4,16" "1,1,4,4"; do
  echo "== tile 12 MNX_DW_SPLIT=$sp"
  MNX_DW_TILE=12,12,12,12 MNX_DW_SPLIT=$sp timeout 300 python tools/quick_cn_bench.py 2>&1 | tail -2
done
echo "== tile 8 split 1,1,2,16"; MNX_DW_TILE=8,8,8,8 MNX_DW_SPLIT=1,1,2,16 timeout 300 python tools/quick_cn_bench.py 2>&1 | tail -2
) | tee gpurun_out/iter_dwsplit12.log
timeout 600 python -m pytest tests/test_gpu_convnext.py -q -m gpu 2>&1 | tail -3
