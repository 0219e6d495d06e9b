#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); s=d['stages']; print(sys.argv[1], round(d['value'],1), 'decode_ms', round(s['decode_ms'],1), 'dec_frac', round(s['decode_frac_hbm'],3))" $1 | tee -a gpurun_out/diag.log; }
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/diag.log; env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline $ARGS > gpurun_out/diag_$name.json 2> gpurun_out/diag_$name.err; summ gpurun_out/diag_$name.json; tail -2 gpurun_out/diag_$name.err; }
ARGS="--arch base --batch 64" run base64_chain A=1
ARGS="--arch base --batch 64" run base64_nochain B200W_NO_CROSS_CHAIN=1
ARGS="--arch turbo --batch 128" run turbo_chain A=1
ARGS="--arch turbo --batch 128" run turbo_nochain B200W_NO_CROSS_CHAIN=1
ARGS="--arch base --batch 256" run base256_chain A=1
ARGS="--arch base --batch 256" run base256_nochain B200W_NO_CROSS_CHAIN=1
