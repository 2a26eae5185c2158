#!/bin/bash
mkdir -p gpurun_out
python tools/gn_bench.py > gpurun_out/gn_plain.log 2>&1 && cat gpurun_out/gn_plain.log | grep "B=" &&
ncu --set full --clock-control none --import-source on -k regex:gn_ -s 40 -c 6 -f -o gpurun_out/prof_gn python tools/gn_bench.py > gpurun_out/ncu_gn.log 2>&1
echo "ncu exit $?"
