#!/bin/bash
# ncu evidence for profiles/: launch lists (gpu__time_duration) and --set full captures of the top kernels.
# usage: bash tools/gpu_profile.sh <tag>      (outputs gpurun_out/<tag>_*.csv / .ncu-rep)
tag=${1:-r1}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${tag}_launches_swin_cluster16.csv python tools/profile_step.py > gpurun_out/profile.log 2>&1
PROFILE_ENCODER=convnext_base PROFILE_ENCODE_ONLY=1 timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${tag}_launches_convnext_encoder.csv python tools/profile_step.py >> gpurun_out/profile.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:decode_mega16 -c 1 -f -o gpurun_out/${tag}_mega16 python tools/profile_step.py >> gpurun_out/profile.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:gemm_tc_kernel -s 60 -c 2 -f -o gpurun_out/${tag}_gemm_tc python tools/profile_step.py >> gpurun_out/profile.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:window_attn -s 10 -c 1 -f -o gpurun_out/${tag}_winattn python tools/profile_step.py >> gpurun_out/profile.log 2>&1
MNX_DECODE_PATH=graph timeout 600 $NCU --set full --import-source on -k regex:attn_kernel -s 41 -c 2 -f -o gpurun_out/${tag}_xattn python tools/profile_step.py >> gpurun_out/profile.log 2>&1
PROFILE_ENCODER=convnext_base PROFILE_ENCODE_ONLY=1 timeout 600 $NCU --set full --import-source on -k regex:dwconv_ln -s 8 -c 1 -f -o gpurun_out/${tag}_dwconv python tools/profile_step.py >> gpurun_out/profile.log 2>&1
tail -3 gpurun_out/profile.log; ls -la gpurun_out/ | tail -12
