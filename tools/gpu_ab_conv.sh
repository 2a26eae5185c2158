#!/bin/bash
# A/B of conv_tc variants on one box: per-launch tables of one guided step with each library build / env switch.
#   bash tools/gpu_ab_conv.sh name1 "ENV=.. ENV=.." name2 "..." ...
mkdir -p gpurun_out
while [ $# -ge 2 ]; do
  name=$1; envs=$2; shift 2
  env $envs python tools/conv_table.py --reps 10 > gpurun_out/ab_$name.txt 2>&1
  echo "$name: $(tail -1 gpurun_out/ab_$name.txt | cut -c1-70)"
done
