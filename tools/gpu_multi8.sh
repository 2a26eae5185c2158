#!/bin/bash
# N-GPU bench lines through torchrun (one rank per GPU, NCCL):   bash tools/gpu_multi8.sh N tag [bench.py args...]
mkdir -p gpurun_out
N=${1:-8}; TAG=${2:-n$N}; shift 2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 3 "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench $TAG (N=$N $@) exit $?"; tail -n 2 gpurun_out/bench_$TAG.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
    print('$TAG', 'N', d['n_gpus'], round(d['value'],2), 'img/s  e2e', round(d['e2e']['value'],2), 'sha', (d.get('images_sha256') or '')[:16], 'strong', d.get('strong_scaling'), 'clocks', d['clocks']['sm_mhz'])
except Exception as e:
    print('$TAG failed', e)
PY
