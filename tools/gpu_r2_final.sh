#!/bin/bash
# round-2 evidence pass: GPU suite, smoke, bench (all configs, both arms), ncu launch lists and --set full captures
mkdir -p gpurun_out
tag=${1:-r2h}
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | tee gpurun_out/gpu.txt
timeout 2400 python -m pytest tests -q -m gpu -x -s 2>&1 | grep -vE "^\s*$" | tail -30 | tee gpurun_out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.log
timeout 600 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/${tag}_bench.json | cut -c1-300
tail -3 gpurun_out/bench.err
for c in c1 c3 c4 c5; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 2> gpurun_out/bench_$c.err | tee gpurun_out/${tag}_bench_$c.json | cut -c1-200
  tail -2 gpurun_out/bench_$c.err
done
timeout 900 python bench.py --impl reference 2> gpurun_out/bench_ref.err | tee gpurun_out/${tag}_bench_ref.json | cut -c1-300
NCU="ncu --clock-control none"
# launch list of the bench command itself (the first 1500 launches: warm-up steps of the pipelined arm)
timeout 900 $NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file gpurun_out/${tag}_launches_bench.csv python bench.py --steps 3 --warmup 3 > gpurun_out/profile.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches_bench.csv | tee gpurun_out/${tag}_launches_bench.md
PROFILE_ENCODE_ONLY=1 timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${tag}_launches_swin_encoder.csv python tools/profile_step.py >> gpurun_out/profile.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches_swin_encoder.csv | tee gpurun_out/${tag}_launches_swin_encoder.md
PROFILE_ENCODER=convnext_base PROFILE_ENCODE_ONLY=1 timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${tag}_launches_convnext_encoder.csv python tools/profile_step.py >> gpurun_out/profile.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches_convnext_encoder.csv | tee gpurun_out/${tag}_launches_convnext_encoder.md
FULL="$NCU --set full --import-source on"
PROFILE_ENCODE_ONLY=1 timeout 600 $FULL -k regex:gemm_tc_kernel -s 60 -c 3 -f -o gpurun_out/${tag}_gemm_tc python tools/profile_step.py >> gpurun_out/profile.log 2>&1
PROFILE_ENCODE_ONLY=1 timeout 600 $FULL -k regex:window_attn -s 10 -c 1 -f -o gpurun_out/${tag}_winattn python tools/profile_step.py >> gpurun_out/profile.log 2>&1
PROFILE_ENCODER=convnext_base PROFILE_ENCODE_ONLY=1 timeout 600 $FULL -k regex:dwconv_stats -s 8 -c 1 -f -o gpurun_out/${tag}_dwconv python tools/profile_step.py >> gpurun_out/profile.log 2>&1
MNX_DECODE_PATH=wide timeout 900 $FULL -k regex:decode_wide -c 1 -f -o gpurun_out/${tag}_wide python tools/quick_dec_bench.py 32 >> gpurun_out/profile.log 2>&1
tail -3 gpurun_out/profile.log; ls -la gpurun_out | grep ${tag}
