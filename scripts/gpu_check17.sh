#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import sys; sys.path.insert(0,'tests'); import util; util.model_root('small')" > /dev/null
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); s=d['stages']; print(sys.argv[1], round(d['value'],1), 'decode_ms', round(s['decode_ms'],1), 'per step', round(s['decode_ms']/228,3), 'roof', round(d['roofline']['frac'],3))" $1 | tee -a gpurun_out/diag.log; }
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/diag.log; env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/diag_$name.json 2> gpurun_out/diag_$name.err; summ gpurun_out/diag_$name.json; tail -2 gpurun_out/diag_$name.err; }
run c3 A=1
run c4 B200W_CROSS_CTAS_PER_SM=4
run c2 B200W_CROSS_CTAS_PER_SM=2
run c4_mb3 B200W_CROSS_CTAS_PER_SM=4 B200W_N_MICROBATCH=3
run c4_nochain B200W_CROSS_CTAS_PER_SM=4 B200W_NO_CROSS_CHAIN=1
B200W_CROSS_CTAS_PER_SM=4 python scripts/trace_decode.py small 256 10 gpurun_out/trace_dec_small256_stream_c4.json 2>&1 | tail -1
