#!/bin/bash
# fused LayerNorm + query projection in the cross-attention kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=600 bash scripts/gpu_bringup.sh tests/test_gpu_decoder.py tests/test_gpu_api.py tests/test_gpu_fullsize.py tests/test_gpu_turbo.py
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['stages'].items()}, round(d['roofline']['frac'],3))" $1 | tee -a gpurun_out/bringup.log; }
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/bringup.log; env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; summ gpurun_out/bench_$name.json; tail -2 gpurun_out/bench_$name.err; }
run v13_fused A=1
run v13_unfused B200W_NO_FUSED_Q=1
run v13_fused_noprio B200W_NO_PRIORITY=1
run v13_fused_mb1 B200W_NO_MICROBATCH=1
