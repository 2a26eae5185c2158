#!/bin/bash
# round-2 ncu evidence: launch lists (one guided step, one decode, the bench) and --set full captures of the dominant
# convolution, the GroupNorm-in-operand-path variant and the wide decoder attention.  Each ncu run follows a plain run
# of the same command that exited 0.
mkdir -p gpurun_out
python tools/profile_step.py --batch 64 --what unet > gpurun_out/plain_unet.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_unet.csv python tools/profile_step.py --batch 64 --what unet > gpurun_out/ncu_unet.log 2>&1
echo "unet launch list exit $?"; tail -1 gpurun_out/plain_unet.log
python tools/profile_step.py --batch 64 --what decode > gpurun_out/plain_decode.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_decode.csv python tools/profile_step.py --batch 64 --what decode > gpurun_out/ncu_decode.log 2>&1
echo "decode launch list exit $?"; tail -1 gpurun_out/plain_decode.log
python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --skip-secondary > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --skip-secondary > gpurun_out/ncu_bench.log 2>&1
echo "bench launch list exit $?"
# full captures
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 17 -c 3 \
    -f -o gpurun_out/r02_prof_conv_tc python tools/profile_step.py --batch 64 --what unet > gpurun_out/ncu_full.log 2>&1
echo "conv_tc full capture exit $?"
STEDM_GN_FUSION=1 python tools/profile_step.py --batch 64 --what unet > gpurun_out/plain_unet_xf.log 2>&1 &&
STEDM_GN_FUSION=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 17 -c 2 \
    -f -o gpurun_out/r02_prof_conv_tc_xf python tools/profile_step.py --batch 64 --what unet > gpurun_out/ncu_full_xf.log 2>&1
echo "conv_tc XF full capture exit $?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attention_wide -c 1 \
    -f -o gpurun_out/r02_prof_attn_wide python tools/profile_step.py --batch 64 --what decode > gpurun_out/ncu_full_attn.log 2>&1
echo "attention_wide full capture exit $?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 2 \
    -f -o gpurun_out/r02_prof_conv_tc_bn128 python tools/profile_step.py --batch 64 --what unet > gpurun_out/ncu_full128.log 2>&1
echo "conv_tc BN=128 full capture exit $?"
ls -la gpurun_out/*.ncu-rep
