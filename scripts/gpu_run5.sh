#!/bin/bash
# round 2, GPU run 5: fused LN (push model) and step-boundary kernel, each A/B'd by env switch on one box; stream kernel for mid-size batches
cd "$(dirname "$0")/.."
O=gpurun_out/run5; mkdir -p $O
./ab_build/mufu_bench | tee $O/mufu.txt
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x -k "7" 2>&1 | tail -3 | tee $O/tests_a.log
timeout 1200 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_eot_compaction.py tests/test_gpu_invariance.py tests/test_gpu_small256.py tests/test_gpu_turbo.py -m gpu -q -x 2>&1 | grep -v Warning | tail -8 | tee $O/tests_b.log
B="python bench.py --no-extras --no-cpu-baseline --steps 3 --warmup 3"
fmt="import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], 'value', round(d['value']), 'dec_ms', round(d['stages']['decode_ms'],1), 'frac', round(d['stages']['decode_frac_hbm'],3), 'clk', d['clocks']['sm_mhz'])"
run() { # name, env..., -- bench args
  name=$1; shift
  env "$@" 2>>$O/err.log | python -c "$fmt" "$name" | tee -a $O/ab.txt
}
for rep in 1 2; do
  for cfg in "small256 --config 2" "base64 --config 1" "small32 --config 2 --batch 32"; do
    set -- $cfg; n=$1; shift
    run "$n ln+bd" A=1 $B "$@"
    run "$n ln" B200W_NO_STEP_BOUNDARY=1 $B "$@"
    run "$n bd" B200W_NO_FUSED_LN=1 $B "$@"
    run "$n none" B200W_NO_FUSED_LN=1 B200W_NO_STEP_BOUNDARY=1 $B "$@"
  done
done
run "turbo128 bd" A=1 $B --config 3
run "turbo128 none" B200W_NO_STEP_BOUNDARY=1 $B --config 3
run "base64 stream" B200W_CROSS_SPLIT=0 B200W_NO_FUSED_LN=1 $B --config 1
run "small32 stream" B200W_CROSS_SPLIT=0 B200W_NO_FUSED_LN=1 $B --config 2 --batch 32
run "small64 split" B200W_NO_FUSED_LN=1 $B --config 2 --batch 64
run "small64 stream" B200W_CROSS_SPLIT=0 B200W_NO_FUSED_LN=1 $B --config 2 --batch 64
tail -3 $O/err.log
