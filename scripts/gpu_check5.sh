#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=200 bash scripts/gpu_bringup.sh tests/test_gpu_decoder.py
timeout 600 python - <<'PY' 2>&1 | tail -8 | tee -a gpurun_out/bringup.log
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, __graft_entry__ as g, util
pkg=g.load_package(); B=256
eng=pkg.Engine(util.model_root('small'),'small',0,B)
pcm=np.stack([util.synth_audio('N',480000,2000+i) for i in range(4)]*64)
eng.upload_pcm(pcm); eng.transcribe_resident(B, max_new_tokens=4, honor_eot=False)
L=12; d=768
for it in range(2):
    ms=eng.time_stage(3,B,8)/(8*L)
    print('cross-attn alone: %.4f ms/launch  %.1f GB/s  frac %.3f'%(ms, B*1500*d*4/ms/1e6, B*1500*d*4/ms/1e6/6534.1))
ms=eng.time_stage(2,B,1,n_steps=228); print('decode 228 steps: %.1f ms  (%.3f ms/step)'%(ms, ms/228))
ms=eng.time_stage(2,B,1,n_steps=228); print('decode 228 steps: %.1f ms  (%.3f ms/step)'%(ms, ms/228))
PY
