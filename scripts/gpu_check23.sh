#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for p in 0 1 2; do echo "== poly $p" | tee -a gpurun_out/bringup.log; B200W_ATTN_POLY=$p T=600 bash scripts/gpu_bringup.sh tests/test_gpu_attention.py tests/test_gpu_encoder.py 2>&1 | grep -E "max\||cross_|passed|failed" ; done
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); s=d['stages']; print(sys.argv[1], round(d['value'],1), 'enc_ms', round(s['encoder_ms'],2), 'enc_frac', round(s['encoder_frac_tensor_sustained'],3), 'decode_ms', round(s['decode_ms'],1))" $1 | tee -a gpurun_out/diag.log; }
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/diag.log; env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --new-tokens 28 > gpurun_out/diag_$name.json 2> gpurun_out/diag_$name.err; summ gpurun_out/diag_$name.json; tail -2 gpurun_out/diag_$name.err; }
run poly0 B200W_ATTN_POLY=0
run poly1 B200W_ATTN_POLY=1
run poly2 B200W_ATTN_POLY=2
run poly0b B200W_ATTN_POLY=0
run poly1b B200W_ATTN_POLY=1
for p in 0 1 2; do echo "=== turbo poly $p" | tee -a gpurun_out/diag.log; B200W_ATTN_POLY=$p timeout 600 python bench.py --arch turbo --batch 128 --steps 2 --warmup 3 --no-cpu-baseline --new-tokens 28 > gpurun_out/diag_turbo_poly$p.json 2> gpurun_out/diag_turbo_poly$p.err; summ gpurun_out/diag_turbo_poly$p.json; done
