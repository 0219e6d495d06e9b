#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # arch B label
timeout 600 python - <<PY 2>&1 | tail -1 | tee -a gpurun_out/bringup.log
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, __graft_entry__ as g, util
pkg=g.load_package(); B=$2
eng=pkg.Engine(util.model_root('$1'),'$1',0,B)
pcm=np.stack([util.synth_audio('N',480000,2000+i) for i in range(4)]*(B//4))
eng.upload_pcm(pcm); eng.transcribe_resident(B, max_new_tokens=4, honor_eot=False)
ms=min(eng.time_stage(2,B,1,n_steps=228) for _ in range(2))
print('$1 B=$2 $3: decode 228 steps %.1f ms (%.3f ms/step)'%(ms, ms/228))
PY
}
echo "=== decode sweeps" | tee -a gpurun_out/bringup.log
for bn in 32 64 128; do B200W_DEC_BN=$bn run small 256 "bn=$bn"; done
B200W_NO_PDL=1 run small 256 "nopdl"
B200W_NO_GRAPH=1 run small 256 "nograph"
B200W_MICROBATCH_MIN=32 run base 64 "mb"
run base 64 "nomb"
B200W_DEC_BN=32 run base 64 "nomb bn=32"
