#!/bin/bash
# Evidence refresh after a late change: full GPU test suite, smoke, bench (with `secondary`), the guided step's launch
# list and the convolution tables.  The long legs (reference arm, sweeps, ncu --set full) stay in gpu_final.sh /
# gpu_r2_profiles.sh.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python tools/profile_step.py --batch 64 --what unet > gpurun_out/plain_unet.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_unet.csv python tools/profile_step.py --batch 64 --what unet > gpurun_out/ncu_unet.log 2>&1
echo "unet launch list exit $?"; tail -1 gpurun_out/plain_unet.log
timeout 300 python tools/conv_table.py --reps 10 > gpurun_out/r02_conv_table_unet_step.txt 2>&1; tail -1 gpurun_out/r02_conv_table_unet_step.txt
timeout 300 python tools/conv_table.py --latent 128 --reps 5 > gpurun_out/r02_conv_table_unet_step_l128.txt 2>&1; tail -1 gpurun_out/r02_conv_table_unet_step_l128.txt
timeout 300 python tools/phase_times.py 2>&1 | tail -2 > gpurun_out/r02_phase_times.txt; cat gpurun_out/r02_phase_times.txt
python - <<PY
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('bench', round(d['value'],2), 'img/s  e2e', round(d['e2e']['value'],2), ' unet step', round(d['unet_step_ms'],2), 'ms  clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'sha', d['images_sha256'][:16], 'roofline', d['roofline']['frac'])
print('secondary', json.dumps(d.get('secondary'))[:600])
PY
