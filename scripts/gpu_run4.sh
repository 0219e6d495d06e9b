#!/bin/bash
# round 2, GPU run 4: fused residual+LayerNorm decoder GEMMs, EOT compaction, whisper_srv; A/B against the unfused path on one box
cd "$(dirname "$0")/.."
O=gpurun_out/run4; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_attention.py -m gpu -q -x 2>&1 | tail -5 | tee $O/tests_a.log
timeout 1200 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_eot_compaction.py tests/test_gpu_srv.py tests/test_gpu_invariance.py tests/test_gpu_parity_configs.py tests/test_gpu_api.py -m gpu -q -x -s 2>&1 | grep -v Warning | tail -40 | tee $O/tests_b.log
B="python bench.py --no-extras --no-cpu-baseline --steps 3 --warmup 3"
fmt="import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'dec_ms', round(d['stages']['decode_ms'],1), 'frac', round(d['stages']['decode_frac_hbm'],3), 'enc_ms', round(d['stages']['encoder_ms'],1), 'clk', d['clocks']['sm_mhz'])"
for rep in 1 2; do
  for mode in fused unfused; do
    if [ $mode = unfused ]; then export B200W_NO_FUSED_LN=1; else unset B200W_NO_FUSED_LN; fi
    $B --config 2 2>>$O/err.log | python -c "$fmt" "small256-$mode" | tee -a $O/ab.txt
    $B --config 1 2>>$O/err.log | python -c "$fmt" "base64-$mode" | tee -a $O/ab.txt
    $B --config 2 --batch 32 2>>$O/err.log | python -c "$fmt" "small32-$mode" | tee -a $O/ab.txt
  done
done
unset B200W_NO_FUSED_LN
for sp in 2 3 6; do
  B200W_CROSS_SPLIT=$sp $B --config 1 2>>$O/err.log | python -c "$fmt" "base64-split$sp" | tee -a $O/ab.txt
  B200W_CROSS_SPLIT=$sp $B --config 2 --batch 32 2>>$O/err.log | python -c "$fmt" "small32-split$sp" | tee -a $O/ab.txt
done
python scripts/trace_decode.py base 64 6 $O/trace_base64.json >> $O/err.log 2>&1
python scripts/trace_decode.py small 32 6 $O/trace_small32.json >> $O/err.log 2>&1
tail -3 $O/err.log
