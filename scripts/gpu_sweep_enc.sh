#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== encoder sub-batch sweep (small, B=128), ms per encoder pass" | tee -a gpurun_out/bringup.log
for sub in 8 16 24 32 64 128; do
B200W_ENC_SUB_BATCH=$sub timeout 300 python - <<PY 2>&1 | tail -1 | tee -a gpurun_out/bringup.log
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, __graft_entry__ as g, util
pkg=g.load_package(); B=128
eng=pkg.Engine(util.model_root('small'),'small',0,B)
pcm=np.stack([util.synth_audio('N',480000,2000+i) for i in range(4)]*32)
eng.upload_pcm(pcm); eng.time_stage(0,B,1); eng.time_stage(1,B,1)
ms=min(eng.time_stage(1,B,2)/2 for _ in range(2))
print('sub=$sub encoder_ms=%.2f  tensor_frac_sustained=%.3f'%(ms, 386.63e9*B/(ms/1e3)/1e12/1385.0))
PY
done
