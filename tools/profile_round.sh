#!/bin/bash
# Profiling pass of one round (run under gpurun, ONE GPU):  bash tools/profile_round.sh r02h
# 1) ncu launch list of one timed step of the default bench command (per-kernel shares; graph replays included),
# 2) --set full captures of the first conv launches of an inference forward (march / slide / tap),
# 3) --set full captures of the weight-gradient and norm-backward kernels of a training step,
# 4) --set full captures of the HBM-bound kernels,
# 5) CUPTI per-kernel tables (warm, unserialised) and per-shape conv tables of both workloads.
# The .ncu-rep files are summarised ON THE BOX (tools/summarize_ncu.py -> markdown) and deleted: gpurun brings back at
# most 64 MiB.
tag=${1:-r02h}
mkdir -p gpurun_out
o=gpurun_out/${tag}
ncu --metrics gpu__time_duration.sum --clock-control none -s 2200 -c 1600 --csv --log-file ${o}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train > ${o}_launches_bench.log 2>&1
python tools/summarize_ncu.py launches ${o}_launches.csv ${o}_launches_v2_tta8.md \
    "ncu launch list: 1600 consecutive launches (> one step) of python bench.py --steps 1 --warmup 1 --no-train" > /dev/null
rm -f ${o}_launches.csv
ncu --set full --clock-control none -k regex:"conv_input|conv_slide|conv_march|conv_tap" -c 9 -f -o ${o}_conv \
    python tools/layer_profile.py v2_tta8 > ${o}_conv.log 2>&1
python tools/summarize_ncu.py full ${o}_conv.ncu-rep ${o}_conv_full.md "ncu --set full: first 8 conv launches of a v2_tta8 window batch (9 x 128^3)" > /dev/null
ncu --set full --clock-control none -k regex:"conv_wgrad|norm_bwd_reduce|norm_bwd_apply" -c 8 -f -o ${o}_train \
    python tools/layer_profile.py v2_train > ${o}_train.log 2>&1
python tools/summarize_ncu.py full ${o}_train.ncu-rep ${o}_train_full.md "ncu --set full: first wgrad / norm-backward launches of a v2_train step" > /dev/null
ncu --set full --clock-control none -k regex:"upsample2x|affine_pool|head_conv|blend_acc|tta_acc|conv_point|pack_windows" -c 10 -f \
    -o ${o}_elem python tools/layer_profile.py v2_tta8 > ${o}_elem.log 2>&1
python tools/summarize_ncu.py full ${o}_elem.ncu-rep ${o}_elem_full.md "ncu --set full: HBM-bound kernels of a v2_tta8 step" > /dev/null
rm -f ${o}_conv.ncu-rep ${o}_train.ncu-rep ${o}_elem.ncu-rep
python tools/kernel_times.py v2_tta8 > ${o}_kernels_v2_tta8.md 2>> ${o}_kt.err
python tools/kernel_times.py v2_train > ${o}_kernels_v2_train.md 2>> ${o}_kt.err
python tools/layer_profile.py v2_train > ${o}_layers_v2_train.md 2>> ${o}_kt.err
python tools/layer_profile.py v2_tta8 > ${o}_layers_v2_tta8.md 2>> ${o}_kt.err
du -sh gpurun_out; ls -la gpurun_out/ | grep ${tag}
