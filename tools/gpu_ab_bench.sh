#!/bin/bash
# same-box A/B of the whole bench: alternate env settings, 2 rounds.   bash tools/gpu_ab_bench.sh "A=1" "A=0" ...
mkdir -p gpurun_out
for round in 1 2; do
  i=0
  for envs in "$@"; do
    i=$((i+1))
    env $envs python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/ab_bench_${i}_${round}.json 2> gpurun_out/ab_bench_${i}_${round}.err
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_bench_${i}_${round}.json').read().strip().splitlines()[-1])
    print('[$envs] round $round:', round(d['value'],2), 'img/s  e2e', round(d['e2e']['value'],2), ' unet step', round(d['unet_step_ms'],2), 'ms  clocks', d['clocks']['sm_mhz'], 'launches', d['gpu_launches'])
except Exception as e:
    print('[$envs] failed', e)
PY
  done
done
