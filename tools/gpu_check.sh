#!/bin/bash
# Run on the GPU box (through gpurun): each stage in its own process so a faulting kernel cannot poison the rest.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in ${@:-ops tc model}; do
  echo "=== $f ===" 
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q -s -x --no-header -p no:cacheprovider > gpurun_out/pytest_$f.log 2>&1
  echo "exit $?" >> gpurun_out/pytest_$f.log
  tail -n 25 gpurun_out/pytest_$f.log
done
