#!/bin/bash
# Profiling pass of one round (run under gpurun, ONE GPU):  bash tools/profile_round.sh r01c
# 1) launch list of one v2_tta8 bench step (shares per kernel), 2) --set full captures of the first launches of the
# elementwise kernels and of the conv kernels in a forward pass.  Summarise here with tools/summarize_ncu.py.
tag=${1:-r01c}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 8400 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline \
    > gpurun_out/${tag}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"norm_apply|scale_pool|upsample2x|se_gate|head_conv|scale_kernel|pack_windows|blend_acc" -c 14 -f \
    -o gpurun_out/${tag}_elem python tools/layer_profile.py v2_tta8 > gpurun_out/${tag}_elem.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv_slide|conv_march|conv_point|conv_tap" -c 9 -f \
    -o gpurun_out/${tag}_conv python tools/layer_profile.py v2_tta8 > gpurun_out/${tag}_conv.log 2>&1
ls -la gpurun_out/
