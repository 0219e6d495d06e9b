#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_fullsize.py tests/test_gpu_small256.py tests/test_gpu_api.py -x -q -p no:cacheprovider 2>&1 | tail -3
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); s=d['stages']; print(sys.argv[1], round(d['value'],1), 'decode_ms', round(s['decode_ms'],1), 'dec_frac', round(s['decode_frac_hbm'],3))" $1 | tee -a gpurun_out/diag.log; }
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/diag.log; env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline $ARGS > gpurun_out/diag_$name.json 2> gpurun_out/diag_$name.err; summ gpurun_out/diag_$name.json; tail -2 gpurun_out/diag_$name.err; }
ARGS="--arch base --batch 64" run base64_auto A=1
ARGS="" run small256_auto A=1
