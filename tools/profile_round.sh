#!/bin/bash
# Profiling pass of one round (run under gpurun, ONE GPU):  bash tools/profile_round.sh r02g
# 1) ncu launch list of the default bench command (per-kernel shares of the v2_tta8 step, graph replays included),
# 2) --set full captures of the first conv launches of an inference forward (march / slide / tap / point),
# 3) --set full captures of the weight-gradient and norm-backward kernels of a training step,
# 4) CUPTI per-kernel tables (warm, unserialised) of both workloads.  Summarise here with tools/summarize_ncu.py.
tag=${1:-r02g}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train \
    > gpurun_out/${tag}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv_slide|conv_march|conv_tap" -c 8 -f \
    -o gpurun_out/${tag}_conv python tools/layer_profile.py v2_tta8 > gpurun_out/${tag}_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv_wgrad|norm_bwd_reduce|norm_bwd_apply" -c 8 -f \
    -o gpurun_out/${tag}_train python tools/layer_profile.py v2_train > gpurun_out/${tag}_train.log 2>&1
ncu --set full --clock-control none -k regex:"upsample2x|affine_pool|head_conv|blend_acc|tta_acc|conv_point" -c 10 -f \
    -o gpurun_out/${tag}_elem python tools/layer_profile.py v2_tta8 > gpurun_out/${tag}_elem.log 2>&1
python tools/kernel_times.py v2_tta8 > gpurun_out/${tag}_kernels_v2_tta8.md 2>> gpurun_out/${tag}_kt.err
python tools/kernel_times.py v2_train > gpurun_out/${tag}_kernels_v2_train.md 2>> gpurun_out/${tag}_kt.err
python tools/layer_profile.py v2_train > gpurun_out/${tag}_layers_v2_train.md 2>> gpurun_out/${tag}_kt.err
python tools/layer_profile.py v2_tta8 > gpurun_out/${tag}_layers_v2_tta8.md 2>> gpurun_out/${tag}_kt.err
ls -la gpurun_out/ | grep ${tag}
