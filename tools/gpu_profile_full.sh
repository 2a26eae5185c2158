#!/bin/bash
# launch list of one guided step + one full-metric capture of the dominant kernel (conv_tc_kernel<256>):
# tensor-core conv launches 17-19 of the step = 1024->1024 3x3 @16x16 on 128 samples (plain, + residual) and the qkv 1x1
mkdir -p gpurun_out
B=${1:-64}
python tools/profile_step.py --batch $B --what unet > gpurun_out/plain_unet.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_unet.csv python tools/profile_step.py --batch $B --what unet > gpurun_out/ncu_unet.log 2>&1
echo "launch list exit $?"; tail -1 gpurun_out/plain_unet.log
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 17 -c 3 \
    -f -o gpurun_out/prof_conv_tc python tools/profile_step.py --batch $B --what unet > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"; ls -la gpurun_out/*.ncu-rep
