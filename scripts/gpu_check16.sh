#!/bin/bash
# streaming (single-wave, rolling-load) cross-attention kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=600 bash scripts/gpu_bringup.sh tests/test_gpu_decoder.py tests/test_gpu_fullsize.py
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); s=d['stages']; print(sys.argv[1], round(d['value'],1), 'decode_ms', round(s['decode_ms'],1), 'per step', round(s['decode_ms']/228,3), 'roof', round(d['roofline']['frac'],3))" $1 | tee -a gpurun_out/diag.log; }
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/diag.log; env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/diag_$name.json 2> gpurun_out/diag_$name.err; summ gpurun_out/diag_$name.json; tail -2 gpurun_out/diag_$name.err; }
run st_mb2 A=1
run old_mb2 B200W_NO_CROSS_STREAM=1
run st_mb2_noprio B200W_NO_PRIORITY=1
run st_mb3 B200W_N_MICROBATCH=3
run st_mb4 B200W_N_MICROBATCH=4
run st_mb1 B200W_NO_MICROBATCH=1
python scripts/trace_decode.py small 256 10 gpurun_out/trace_dec_small256_stream.json 2>&1 | tail -1
