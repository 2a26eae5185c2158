#!/bin/bash
# End-of-round evidence refresh: full GPU test suite, smoke, bench (256^2 and 512^2), reference arm, predict tail bench,
# launch lists and the ncu --set full capture of the dominant kernel.
mkdir -p gpurun_out
bash tools/gpu_check.sh ops tc style model
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
python bench.py --latent 128 --batch 64 --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_l128.json 2> gpurun_out/bench_l128.err; echo "bench 512 exit $?"
python bench.py --n-style 10 --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_n10.json 2> gpurun_out/bench_n10.err; echo "bench N=10 exit $?"
timeout 300 python tools/predict_bench.py > gpurun_out/predict_bench.txt 2>&1; cat gpurun_out/predict_bench.txt | tail -3
timeout 200 python tools/style_bench.py --images 64 640 2>&1 | grep -v fp32 > gpurun_out/style_bench.txt
bash tools/gpu_launchlist_style.sh 256 > /dev/null
bash tools/gpu_profile_full.sh 64 > gpurun_out/profile_full.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_unet.csv > gpurun_out/launches_unet.txt
bash tools/gpu_launchlist_decode.sh 64 > /dev/null; python tools/summarize_launches.py gpurun_out/launches_decode.csv > gpurun_out/launches_decode.txt
for f in bench bench_l128 bench_n10; do python - <<PY
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f', round(d['value'],2), 'img/s  e2e', round(d['e2e']['value'],2), ' unet step', round(d['unet_step_ms'],2), 'ms  clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
done
