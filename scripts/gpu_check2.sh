#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=300 bash scripts/gpu_bringup.sh tests/test_gpu_gemm.py tests/test_gpu_encoder.py tests/test_gpu_decoder.py
echo "=== bench small b256" | tee -a gpurun_out/bringup.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small256_v2.json 2> gpurun_out/bench_small256_v2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_small256_v2.json')); print(d['value'], d['stages'], d['roofline']['frac'])" | tee -a gpurun_out/bringup.log; tail -3 gpurun_out/bench_small256_v2.err
