#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=600 bash scripts/gpu_bringup.sh tests/test_gpu_decoder.py tests/test_gpu_api.py tests/test_gpu_fullsize.py
echo "=== bench small b256" | tee -a gpurun_out/bringup.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small256_v6.json 2> gpurun_out/bench_small256_v6.err; python -c "
import json; d=json.load(open('gpurun_out/bench_small256_v6.json')); print(d['value'], d['e2e']['value'], d['stages'], d['roofline']['frac'])" | tee -a gpurun_out/bringup.log; tail -3 gpurun_out/bench_small256_v6.err
echo "=== bench base b64" | tee -a gpurun_out/bringup.log
timeout 600 python bench.py --arch base --batch 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_base64_v6.json 2> gpurun_out/bench_base64_v6.err; python -c "
import json; d=json.load(open('gpurun_out/bench_base64_v6.json')); print(d['value'], d['stages'], d['roofline']['frac'])" | tee -a gpurun_out/bringup.log
