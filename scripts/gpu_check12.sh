#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=600 bash scripts/gpu_bringup.sh tests/test_gpu_attention.py tests/test_gpu_encoder.py
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['stages'].items()}, round(d['roofline']['frac'],3))" $1 | tee -a gpurun_out/bringup.log; }
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/bringup.log; env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; summ gpurun_out/bench_$name.json; tail -2 gpurun_out/bench_$name.err; }
run v12 A=1
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "enc/" -k regex:encoder_attention -c 1 -o gpurun_out/prof_attn_v2 -f python scripts/profile_kernels.py small 128 6 > gpurun_out/prof_attn_v2.log 2>&1; tail -2 gpurun_out/prof_attn_v2.log
