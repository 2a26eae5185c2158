#!/bin/bash
# A/B of conv_tc builds on one box: per-launch tables of one guided step with each library / switch.
mkdir -p gpurun_out
run() {  # name, env...
  local name=$1; shift
  env "$@" python tools/conv_table.py --reps 10 > gpurun_out/ab_$name.txt 2>&1
  echo "$name: $(tail -1 gpurun_out/ab_$name.txt | cut -c1-70)"
}
[ -f stedm_b200/libstedm_old.so ] && run old STEDM_B200_LIB=$PWD/stedm_b200/libstedm_old.so
run new_nohalo STEDM_TC_HALO=0
run new_halo STEDM_TC_HALO=1
[ -f stedm_b200/libstedm_old.so ] && run old2 STEDM_B200_LIB=$PWD/stedm_b200/libstedm_old.so
run new_halo2 STEDM_TC_HALO=1
