#!/bin/bash
# A/B: shared-memory carve-out of the decode kernels (co-residency with GEMM CTAs), attention kernel with pair barriers
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=600 bash scripts/gpu_bringup.sh tests/test_gpu_attention.py tests/test_gpu_encoder.py tests/test_gpu_decoder.py
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['stages'].items()}, round(d['roofline']['frac'],3))" $1 | tee -a gpurun_out/bringup.log; }
for cfg in "carve prio" "carve noprio" "nocarve prio"; do
  set -- $cfg
  if [ $1 = nocarve ]; then export B200W_NO_CARVEOUT=1; else unset B200W_NO_CARVEOUT; fi
  if [ $2 = noprio ]; then export B200W_NO_PRIORITY=1; else unset B200W_NO_PRIORITY; fi
  echo "=== bench small b256 $1 $2" | tee -a gpurun_out/bringup.log
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small256_$1_$2.json 2> gpurun_out/bench_small256_$1_$2.err; summ gpurun_out/bench_small256_$1_$2.json; tail -2 gpurun_out/bench_small256_$1_$2.err
done
unset B200W_NO_CARVEOUT B200W_NO_PRIORITY
for n in 3 4; do
  export B200W_N_MICROBATCH=$n
  echo "=== bench small b256 carve prio n_mb=$n" | tee -a gpurun_out/bringup.log
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small256_carve_mb$n.json 2> gpurun_out/bench_small256_carve_mb$n.err; summ gpurun_out/bench_small256_carve_mb$n.json
done
