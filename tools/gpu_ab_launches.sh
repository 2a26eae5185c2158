#!/bin/bash
# same-box A/B of one guided step's launch list (ncu, cold-cache per-kernel durations):  bash tools/gpu_ab_launches.sh "A=1" "A=0"
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs python tools/profile_step.py --batch 64 --what unet > gpurun_out/ab_plain_$i.log 2>&1 &&
  env $envs ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/ab_launches_$i.csv python tools/profile_step.py --batch 64 --what unet > gpurun_out/ab_ncu_$i.log 2>&1
  echo "[$envs] exit $?  $(tail -1 gpurun_out/ab_plain_$i.log)"
  python tools/summarize_launches.py gpurun_out/ab_launches_$i.csv | head -12
done
