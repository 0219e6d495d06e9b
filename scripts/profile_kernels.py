#!/usr/bin/env python3
"""Workload for ncu: runs the stages inside NVTX ranges so that `ncu --nvtx --nvtx-include "enc/"` (or "dec/", "mel/")
selects the kernels of one stage.  Usage: profile_kernels.py ARCH BATCH [DECODE_STEPS]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import __graft_entry__ as g
import util

arch, B = sys.argv[1], int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
pkg = g.load_package()
eng = pkg.Engine(util.model_root(arch), arch, 0, B)
pcm = np.stack([util.synth_audio("N", 480000, 2000 + i) for i in range(min(B, 8))])
pcm = np.concatenate([pcm] * ((B + len(pcm) - 1) // len(pcm)))[:B]
eng.upload_pcm(pcm)
eng.transcribe_resident(B, max_new_tokens=4, honor_eot=False)  # warm: attributes, graph capture, caches populated
torch.cuda.synchronize()
nvtx = torch.cuda.nvtx
nvtx.range_push("mel")
eng.time_stage(0, B, 1)
nvtx.range_pop()
nvtx.range_push("enc")
eng.time_stage(1, B, 1)
nvtx.range_pop()
nvtx.range_push("xattn")  # the roofline kernel of bench.py: one cross-attention launch per decoder layer over the whole batch
eng.time_stage(3, B, 1)
nvtx.range_pop()
nvtx.range_push("dec")
eng.time_stage(2, B, 1, n_steps=steps)
nvtx.range_pop()
torch.cuda.synchronize()
print("done")
