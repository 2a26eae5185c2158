#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/conv_table.py --reps 10 > gpurun_out/r02_conv_table_unet_step.txt 2>&1; tail -1 gpurun_out/r02_conv_table_unet_step.txt
timeout 300 python tools/conv_table.py --latent 128 --reps 5 > gpurun_out/r02_conv_table_unet_step_l128.txt 2>&1; tail -1 gpurun_out/r02_conv_table_unet_step_l128.txt
timeout 300 python tools/conv_table.py --what decode --reps 5 > gpurun_out/r02_conv_table_decode.txt 2>&1; tail -1 gpurun_out/r02_conv_table_decode.txt
timeout 200 python tools/attn_bench.py > gpurun_out/r02_attention_wide.txt 2>&1; cat gpurun_out/r02_attention_wide.txt
timeout 200 python tools/xf_bench.py --reps 30 > gpurun_out/r02_xf_bench.txt 2>&1; cat gpurun_out/r02_xf_bench.txt
timeout 900 python tests/perf_sweep.py --batches 1 2 4 8 16 64 256 > gpurun_out/r02_sweep_unet_eps_step.txt 2>&1; cat gpurun_out/r02_sweep_unet_eps_step.txt
timeout 300 python tools/phase_times.py 2>&1 | tail -2 > gpurun_out/r02_phase_times.txt; cat gpurun_out/r02_phase_times.txt
