#!/bin/bash
# attention v2 (two query tiles per CTA, O in TMEM) + thin decode GEMM CTAs (72 regs, red.add residual) with launch priorities
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=600 bash scripts/gpu_bringup.sh tests/test_gpu_gemm.py tests/test_gpu_attention.py tests/test_gpu_encoder.py tests/test_gpu_decoder.py
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['stages'].items()}, round(d['roofline']['frac'],3))" $1 | tee -a gpurun_out/bringup.log; }
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/bringup.log; env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; summ gpurun_out/bench_$name.json; tail -2 gpurun_out/bench_$name.err; }
run v11_prio A=1
run v11_noprio B200W_NO_PRIORITY=1
run v11_prio_nochain B200W_NO_CROSS_CHAIN=1
run v11_prio_mb4 B200W_N_MICROBATCH=4
