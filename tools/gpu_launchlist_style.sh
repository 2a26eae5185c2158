#!/bin/bash
# Launch list of the native style encoder on B style images (default 256 = one chunk at 256x256).
mkdir -p gpurun_out
B=${1:-256}
python tools/profile_step.py --batch $B --what style > gpurun_out/plain_style.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_style.csv python tools/profile_step.py --batch $B --what style > gpurun_out/ncu_style.log 2>&1
echo "launch list exit $?"; tail -1 gpurun_out/plain_style.log
python tools/summarize_launches.py gpurun_out/launches_style.csv | tee gpurun_out/launches_style.txt
