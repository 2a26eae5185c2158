#!/bin/bash
# launch list (per-kernel device time) of one guided U-Net step and one VQ decode at the bench workload
mkdir -p gpurun_out
B=${1:-64}
python tools/profile_step.py --batch $B --what unet > gpurun_out/plain_unet.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_unet.csv python tools/profile_step.py --batch $B --what unet > gpurun_out/ncu_unet.log 2>&1
echo "unet ncu exit $?"; cat gpurun_out/plain_unet.log | tail -1
python tools/profile_step.py --batch $B --what decode > gpurun_out/plain_decode.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_decode.csv python tools/profile_step.py --batch $B --what decode > gpurun_out/ncu_decode.log 2>&1
echo "decode ncu exit $?"; cat gpurun_out/plain_decode.log | tail -1
