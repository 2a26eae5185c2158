#!/bin/bash
# ncu --set full of the first four token GEMMs of the style encoder (stage 1: qkv, proj, fc1+GELU, fc2) on B images
mkdir -p gpurun_out
B=${1:-256}
python tools/profile_step.py --batch $B --what style > gpurun_out/plain_style.log 2>&1 &&
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 4 \
    -f -o gpurun_out/prof_style_gemm python tools/profile_step.py --batch $B --what style > gpurun_out/ncu_style_full.log 2>&1
echo "full capture exit $?"; ls -la gpurun_out/prof_style_gemm.ncu-rep
