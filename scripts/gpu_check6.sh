#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=300 bash scripts/gpu_bringup.sh tests/test_gpu_gemm.py tests/test_gpu_encoder.py tests/test_gpu_decoder.py
echo "=== encoder time small B=64: 2-CTA vs 1-CTA GEMM" | tee -a gpurun_out/bringup.log
for mode in 2cta 1cta; do
if [ $mode = 1cta ]; then export B200W_GEMM_1CTA=1; else unset B200W_GEMM_1CTA; fi
timeout 300 python - <<PY 2>&1 | tail -1 | tee -a gpurun_out/bringup.log
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, __graft_entry__ as g, util
pkg=g.load_package(); B=64
eng=pkg.Engine(util.model_root('small'),'small',0,B)
pcm=np.stack([util.synth_audio('N',480000,2000+i) for i in range(4)]*16)
eng.upload_pcm(pcm); eng.time_stage(0,B,1); eng.time_stage(1,B,1)
ms=eng.time_stage(1,B,3)/3
print('gemm=$mode encoder_ms=%.2f  tensor_frac_sustained=%.3f'%(ms, 386.63e9*B/(ms/1e3)/1e12/1385.0))
PY
done
unset B200W_GEMM_1CTA
echo "=== bench small b256" | tee -a gpurun_out/bringup.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small256_v5.json 2> gpurun_out/bench_small256_v5.err; python -c "
import json; d=json.load(open('gpurun_out/bench_small256_v5.json')); print(d['value'], d['e2e']['value'], d['stages'], d['roofline']['frac'])" | tee -a gpurun_out/bringup.log; tail -3 gpurun_out/bench_small256_v5.err
