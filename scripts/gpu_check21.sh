#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=600 bash scripts/gpu_bringup.sh tests/test_gpu_attention.py tests/test_gpu_decoder.py
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); s=d['stages']; print(sys.argv[1], round(d['value'],1), 'decode_ms', round(s['decode_ms'],1), 'per step', round(s['decode_ms']/228,3), 'roof', round(d['roofline']['frac'],3))" $1 | tee -a gpurun_out/diag.log; }
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/diag.log; env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/diag_$name.json 2> gpurun_out/diag_$name.err; summ gpurun_out/diag_$name.json; tail -2 gpurun_out/diag_$name.err; }
run ef_k8 A=1
run ef_k8_mb3 B200W_N_MICROBATCH=3
run ef_k8_mb4 B200W_N_MICROBATCH=4
timeout 300 ncu --set full --clock-control none --nvtx --nvtx-include "dec/" -k regex:gemm_tcgen05 -s 60 -c 12 -o gpurun_out/prof_decgemm_ef -f python scripts/profile_kernels.py small 256 6 > gpurun_out/prof_decgemm_ef.log 2>&1; tail -1 gpurun_out/prof_decgemm_ef.log
