#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=300 bash scripts/gpu_bringup.sh tests/test_gpu_decoder.py tests/test_gpu_encoder.py
for mode in mb nomb; do
echo "=== bench small b256 $mode" | tee -a gpurun_out/bringup.log
if [ $mode = nomb ]; then export B200W_NO_MICROBATCH=1; else unset B200W_NO_MICROBATCH; fi
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small256_$mode.json 2> gpurun_out/bench_small256_$mode.err; python -c "
import json; d=json.load(open('gpurun_out/bench_small256_$mode.json')); print(d['value'], d['e2e']['value'], d['stages'], d['roofline']['frac'])" | tee -a gpurun_out/bringup.log; tail -3 gpurun_out/bench_small256_$mode.err
done
unset B200W_NO_MICROBATCH
echo "=== bench base b64" | tee -a gpurun_out/bringup.log
timeout 600 python bench.py --arch base --batch 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_base64_v4.json 2> gpurun_out/bench_base64_v4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_base64_v4.json')); print(d['value'], d['stages'], d['roofline']['frac'])" | tee -a gpurun_out/bringup.log
