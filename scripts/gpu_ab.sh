#!/bin/bash
# Same-box A/B of decode-path switches (how every profiles/r02_ab_*.txt was produced): each line = one bench.py run (3 timed
# steps) under the given environment.  usage (on the GPU box): scripts/gpu_ab.sh OUT.txt "NAME ENV=.. ENV=.. -- bench args" ...
#   scripts/gpu_ab.sh gpurun_out/ab.txt "base64-on -- --config 1" "base64-off B200W_NO_STEP_BOUNDARY=1 -- --config 1"
cd "$(dirname "$0")/.."
OUT=$1; shift
B="python bench.py --no-extras --no-cpu-baseline --steps 3 --warmup 3"
fmt="import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], 'value', round(d['value']), 'dec_ms', round(d['stages']['decode_ms'],1), 'frac', round(d['stages']['decode_frac_hbm'],3), 'enc_ms', round(d['stages']['encoder_ms'],1), 'mel_ms', round(d['stages']['mel_ms'],2), 'clk', d['clocks']['sm_mhz'])"
for rep in 1 2; do
  for spec in "$@"; do
    name=${spec%% *}; rest=${spec#* }
    envs=${rest%%--*}; args=${rest#*--}
    env $envs $B $args 2>>$OUT.err | python -c "$fmt" "$name" | tee -a $OUT
  done
done
