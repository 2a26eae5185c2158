#!/bin/bash
mkdir -p gpurun_out
B=${1:-64}
python tools/profile_step.py --batch $B --what decode > gpurun_out/plain_decode.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_decode.csv python tools/profile_step.py --batch $B --what decode > gpurun_out/ncu_decode.log 2>&1
echo "decode launch list exit $?"; tail -1 gpurun_out/plain_decode.log
