#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import sys; sys.path.insert(0,'tests'); import util; util.model_root('small')" > /dev/null
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); s=d['stages']; print(sys.argv[1], round(d['value'],1), 'decode_ms', round(s['decode_ms'],1), 'per step', round(s['decode_ms']/228,3), 'roof', round(d['roofline']['frac'],3))" $1 | tee -a gpurun_out/diag.log; }
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/diag.log; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/diag_$name.json 2> gpurun_out/diag_$name.err; summ gpurun_out/diag_$name.json; tail -2 gpurun_out/diag_$name.err; }
run ef A=1
run noef B200W_NO_EVICT_FIRST=1
run ef2 A=1
run noef2 B200W_NO_EVICT_FIRST=1
run ef_k4 B200W_GRAPH_STEPS=4
run ef_k1 B200W_GRAPH_STEPS=1
