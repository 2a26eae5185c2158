#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${T:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; tail -n ${TAIL:-6} gpurun_out/$name.log; }
T=600 run r2b_tc_gn python -m pytest tests/test_gpu_tc.py -m gpu -q -x --no-header -p no:cacheprovider -k "groupnorm_in_operand"
T=600 TAIL=12 run r2b_model python -m pytest tests/test_gpu_model.py -m gpu -q -s --no-header -p no:cacheprovider -k "groupnorm_in_operand or ancestral or bench_geometry"
T=300 TAIL=60 run r2b_table_fused python tools/conv_table.py --reps 10
python bench.py --steps 3 --warmup 3 --skip-cpu-baseline > gpurun_out/r2b_bench_fused.json 2> gpurun_out/r2b_bench_fused.err; echo "bench fused exit $?"
for f in r2b_bench_fused; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value'],2), 'img/s  e2e', round(d['e2e']['value'],2), ' unet step', round(d['unet_step_ms'],2), 'ms  clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'launches', d['gpu_launches'])
except Exception as e:
    print('$f', 'failed', e)
PY
done
