#!/bin/bash
# sweep of the decode-step micro-batch count / cross-attention hand-over
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B200W_N_MICROBATCH=4 T=600 bash scripts/gpu_bringup.sh tests/test_gpu_decoder.py tests/test_gpu_fullsize.py
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['stages'].items()}, round(d['roofline']['frac'],3))" $1 | tee -a gpurun_out/bringup.log; }
for cfg in "2 0" "3 0" "4 0" "2 1" "3 1" "4 1"; do
  set -- $cfg
  export B200W_N_MICROBATCH=$1
  if [ $2 = 1 ]; then export B200W_NO_CROSS_CHAIN=1; else unset B200W_NO_CROSS_CHAIN; fi
  echo "=== bench small b256 n_mb=$1 nochain=$2" | tee -a gpurun_out/bringup.log
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small256_mb$1_$2.json 2> gpurun_out/bench_small256_mb$1_$2.err; summ gpurun_out/bench_small256_mb$1_$2.json; tail -2 gpurun_out/bench_small256_mb$1_$2.err
done
