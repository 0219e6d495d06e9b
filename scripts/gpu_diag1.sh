#!/bin/bash
# decode-step timing diagnostics (B200W_DIAG_SKIP_CROSS gives wrong tokens on purpose: chain-only time)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import sys; sys.path.insert(0,'tests'); import util; util.model_root('small')" > /dev/null
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); s=d['stages']; print(sys.argv[1], round(d['value'],1), 'decode_ms', round(s['decode_ms'],1), 'per step', round(s['decode_ms']/228,3))" $1 | tee -a gpurun_out/diag.log; }
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/diag.log; env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/diag_$name.json 2> gpurun_out/diag_$name.err; summ gpurun_out/diag_$name.json; }
run mb2 A=1
run mb1 B200W_NO_MICROBATCH=1
run mb1_skipcross B200W_NO_MICROBATCH=1 B200W_DIAG_SKIP_CROSS=1
run mb2_skipcross B200W_DIAG_SKIP_CROSS=1
run mb2_nochain_skipcross B200W_DIAG_SKIP_CROSS=1 B200W_NO_CROSS_CHAIN=1
run mb1_skipcross_nopdl B200W_NO_MICROBATCH=1 B200W_DIAG_SKIP_CROSS=1 B200W_NO_PDL=1
run mb1_skipcross_bn64 B200W_NO_MICROBATCH=1 B200W_DIAG_SKIP_CROSS=1 B200W_DEC_BN=64
