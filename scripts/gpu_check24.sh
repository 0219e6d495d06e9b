#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=600 bash scripts/gpu_bringup.sh tests/test_gpu_gemm.py tests/test_gpu_encoder.py tests/test_gpu_decoder.py tests/test_gpu_fullsize.py 2>&1 | grep -E "passed|failed|error|Error" 
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); s=d['stages']; print(sys.argv[1], round(d['value'],1), 'enc_ms', round(s['encoder_ms'],1), 'decode_ms', round(s['decode_ms'],1), 'dec_frac', round(s['decode_frac_hbm'],3))" $1 | tee -a gpurun_out/diag.log; }
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/diag.log; env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline $ARGS > gpurun_out/diag_$name.json 2> gpurun_out/diag_$name.err; summ gpurun_out/diag_$name.json; tail -2 gpurun_out/diag_$name.err; }
ARGS="" run small_pre A=1
ARGS="--arch turbo --batch 128" run turbo_mb2 A=1
ARGS="--arch turbo --batch 128" run turbo_mb1 B200W_NO_MICROBATCH=1
ARGS="--arch turbo --batch 128" run turbo_mb3 B200W_N_MICROBATCH=3
ARGS="--arch base --batch 64" run base64_mb2 A=1
ARGS="--arch base --batch 64" run base64_mb1 B200W_NO_MICROBATCH=1
ARGS="--arch tiny --batch 64" run tiny64_mb2 A=1
ARGS="--arch tiny --batch 64" run tiny64_mb1 B200W_NO_MICROBATCH=1
