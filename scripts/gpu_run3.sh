#!/bin/bash
# round 2, GPU run 3: same-box A/B (r01 library vs HEAD) on the headline, micro-batch sweeps on the mid-size configs, timelines
cd "$(dirname "$0")/.."
O=gpurun_out/run3; mkdir -p $O
B="python bench.py --no-extras --no-cpu-baseline --steps 3 --warmup 3"
for lib in r01 head r01 head; do
  if [ $lib = r01 ]; then export B200W_LIB=$PWD/ab_build/r01/libax_whisper.so; else unset B200W_LIB; fi
  $B --config 2 2>>$O/err.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$lib', 'small256 value', round(d['value']), 'dec_ms', round(d['stages']['decode_ms'],1), 'enc_ms', round(d['stages']['encoder_ms'],1), 'xattn', d['roofline'].get('ms_per_launch'), d['roofline'].get('whole_batch_launch',{}).get('ms_per_launch'), 'clk', d['clocks']['sm_mhz'])" | tee -a $O/ab.txt
done
unset B200W_LIB
for cfg in 1 3; do for mb in 1 2 4; do
  B200W_N_MICROBATCH=$mb $B --config $cfg 2>>$O/err.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('config$cfg n_mb=$mb value', round(d['value']), 'dec_ms', round(d['stages']['decode_ms'],1), 'frac', round(d['stages']['decode_frac_hbm'],3))" | tee -a $O/mb.txt
done; done
for b in 32 64; do for mb in 1 2 4; do
  B200W_N_MICROBATCH=$mb $B --config 2 --batch $b 2>>$O/err.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('small b=$b n_mb=$mb value', round(d['value']), 'dec_ms', round(d['stages']['decode_ms'],1), 'frac', round(d['stages']['decode_frac_hbm'],3))" | tee -a $O/mb.txt
done; done
python scripts/trace_decode.py base 64 6 $O/trace_base64.json >> $O/err.log 2>&1
python scripts/trace_decode.py small 32 6 $O/trace_small32.json >> $O/err.log 2>&1
python scripts/trace_decode.py small 256 6 $O/trace_small256.json >> $O/err.log 2>&1
python scripts/trace_decode.py turbo 128 6 $O/trace_turbo128.json >> $O/err.log 2>&1
tail -3 $O/err.log
