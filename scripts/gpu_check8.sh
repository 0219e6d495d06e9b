#!/bin/bash
# A/B of the launch-priority change on the decode step, plus decoder tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=600 bash scripts/gpu_bringup.sh tests/test_gpu_decoder.py tests/test_gpu_api.py
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['stages'].items()}, round(d['roofline']['frac'],3))" $1 | tee -a gpurun_out/bringup.log; }
for mode in prio noprio; do
  if [ $mode = noprio ]; then export B200W_NO_PRIORITY=1; else unset B200W_NO_PRIORITY; fi
  echo "=== bench small b256 $mode" | tee -a gpurun_out/bringup.log
  timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small256_$mode.json 2> gpurun_out/bench_small256_$mode.err; summ gpurun_out/bench_small256_$mode.json; tail -2 gpurun_out/bench_small256_$mode.err
done
unset B200W_NO_PRIORITY
echo "=== bench turbo b128" | tee -a gpurun_out/bringup.log
timeout 900 python bench.py --arch turbo --batch 128 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_turbo128_prio.json 2> gpurun_out/bench_turbo128_prio.err; summ gpurun_out/bench_turbo128_prio.json
echo "=== bench base b64" | tee -a gpurun_out/bringup.log
timeout 600 python bench.py --arch base --batch 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_base64_prio.json 2> gpurun_out/bench_base64_prio.err; summ gpurun_out/bench_base64_prio.json
