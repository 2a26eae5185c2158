#!/bin/bash
# round 2, first GPU pass: the GroupNorm-in-operand-path kernel (tests, per-launch table, A/B bench) + the new parity pins
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; timeout ${T:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; tail -n ${TAIL:-6} gpurun_out/$name.log; }
T=600 run r2a_tc_gn python -m pytest tests/test_gpu_tc.py -m gpu -q -x --no-header -p no:cacheprovider -k "groupnorm_in_operand"
T=900 run r2a_tc python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider
T=1500 TAIL=15 run r2a_model python -m pytest tests/test_gpu_model.py -m gpu -q -s --no-header -p no:cacheprovider
T=300 TAIL=3 run r2a_table_fused python tools/conv_table.py --reps 10
T=300 TAIL=3 STEDM_GN_FUSION=0 run r2a_table_unfused python tools/conv_table.py --reps 10
python bench.py --steps 3 --warmup 3 --skip-cpu-baseline > gpurun_out/r2a_bench_fused.json 2> gpurun_out/r2a_bench_fused.err; echo "bench fused exit $?"
STEDM_GN_FUSION=0 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline > gpurun_out/r2a_bench_unfused.json 2> gpurun_out/r2a_bench_unfused.err; echo "bench unfused exit $?"
for f in r2a_bench_fused r2a_bench_unfused; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value'],2), 'img/s  e2e', round(d['e2e']['value'],2), ' unet step', round(d['unet_step_ms'],2), 'ms  clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'launches', d['gpu_launches'])
except Exception as e:
    print('$f', 'failed', e)
PY
done
