#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mel.py tests/test_gpu_encoder.py -x -q -p no:cacheprovider 2>&1 | tail -3
python - <<'PY'
import sys, os
sys.path.insert(0, "tests")
import numpy as np, torch
import __graft_entry__ as g, util
pkg = g.load_package()
eng = pkg.Engine(util.model_root("small"), "small", 0, 256)
pcm = np.stack([util.synth_audio("N", 480000, 2000 + i) for i in range(8)])
pcm = np.concatenate([pcm] * 32)[:256]
eng.upload_pcm(pcm)
eng.time_stage(0, 256, 2)
print("mel ms per 256 chunks:", [round(eng.time_stage(0, 256, 5) / 5, 3) for _ in range(3)])
PY
