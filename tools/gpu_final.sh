#!/bin/bash
# End-of-round evidence refresh (round 2): full GPU test suite, smoke, bench with the driver's flags + reference arm,
# per-launch tables, sweep, attention / style / predict-tail benches.  ncu captures: tools/gpu_r2_profiles.sh;
# multi-GPU lines: tools/gpu_multi8.sh N tag [args] under `gpurun --gpus N`.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
bash tools/gpu_r2_tables.sh
timeout 300 python tools/predict_bench.py > gpurun_out/predict_bench.txt 2>&1; tail -3 gpurun_out/predict_bench.txt
timeout 200 python tools/style_bench.py --images 64 640 2>&1 | grep -v fp32 > gpurun_out/style_bench.txt
python - <<PY
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('bench', round(d['value'],2), 'img/s  e2e', round(d['e2e']['value'],2), ' unet step', round(d['unet_step_ms'],2), 'ms  clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'sha', d['images_sha256'][:16])
PY
