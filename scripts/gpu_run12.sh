#!/bin/bash
# round 2, GPU run 12: thin-A operand boxes + deep pipeline for the decoder-step GEMMs, A/B by env switch on one box
cd "$(dirname "$0")/.."
O=gpurun_out/run12; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_decoder.py tests/test_gpu_invariance.py tests/test_gpu_eot_compaction.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | grep -v Warning | tail -4 | tee $O/tests.log
B="python bench.py --no-extras --no-cpu-baseline --steps 3 --warmup 3"
fmt="import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], 'value', round(d['value']), 'dec_ms', round(d['stages']['decode_ms'],1), 'frac', round(d['stages']['decode_frac_hbm'],3), 'clk', d['clocks']['sm_mhz'])"
run() { name=$1; shift; env "$@" 2>>$O/err.log | python -c "$fmt" "$name" | tee -a $O/ab.txt; }
for rep in 1 2; do
  for cfg in "base64 --config 1" "small32 --config 2 --batch 32" "small64 --config 2 --batch 64" "small16 --config 2 --batch 16 " "turbo128 --config 3" "small256 --config 2"; do
    set -- $cfg; n=$1; shift
    run "$n thin" A=1 $B "$@"
    run "$n full" B200W_NO_THIN_A=1 $B "$@"
  done
done
tail -3 $O/err.log
