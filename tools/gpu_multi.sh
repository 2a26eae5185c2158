#!/bin/bash
# N-GPU bench through torchrun + the reference arm
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit $?"; tail -n 2 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "reference arm exit $?"; cat gpurun_out/bench_ref.json
