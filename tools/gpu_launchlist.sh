#!/bin/bash
mkdir -p gpurun_out
B=${1:-64}
python tools/profile_step.py --batch $B --what unet > gpurun_out/plain_unet.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_unet.csv python tools/profile_step.py --batch $B --what unet > gpurun_out/ncu_unet.log 2>&1
echo "launch list exit $?"; tail -1 gpurun_out/plain_unet.log
