#!/bin/bash
# smoke + default bench on the GPU box; logs under gpurun_out/
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 5 gpurun_out/smoke.log
timeout 1200 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -n 3 gpurun_out/bench.err; cat gpurun_out/bench.json
