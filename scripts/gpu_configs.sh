#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=900 bash scripts/gpu_bringup.sh tests/test_gpu_turbo.py
for cfg in "tiny 64" "base 64" "turbo 128"; do set -- $cfg
echo "=== bench $1 b$2" | tee -a gpurun_out/bringup.log
timeout 900 python bench.py --arch $1 --batch $2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$1_$2.json')); print(d['value'], d['e2e']['value'], d['stages'], d['roofline']['frac'])" | tee -a gpurun_out/bringup.log; tail -2 gpurun_out/bench_$1_$2.err
done
