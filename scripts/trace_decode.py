#!/usr/bin/env python3
"""Kernel timeline of a few decoder steps (CUPTI through torch.profiler): which kernels of the two micro-batches overlap.
Usage: trace_decode.py ARCH BATCH STEPS OUT.json   (writes a compact list: name, stream, start_us, dur_us)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

import __graft_entry__ as g
import util

arch, B, steps, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
pkg = g.load_package()
eng = pkg.Engine(util.model_root(arch), arch, 0, B)
pcm = np.stack([util.synth_audio("N", 480000, 2000 + i) for i in range(min(B, 8))])
pcm = np.concatenate([pcm] * ((B + len(pcm) - 1) // len(pcm)))[:B]
eng.upload_pcm(pcm)
eng.transcribe_resident(B, max_new_tokens=4, honor_eot=False)
eng.time_stage(2, B, 1, n_steps=steps)  # warm graph for this step count
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.time_stage(2, B, 1, n_steps=steps)
    torch.cuda.synchronize()
tmp = out + ".chrome.json"
prof.export_chrome_trace(tmp)
ev = json.load(open(tmp))["traceEvents"]
rows = [(e["name"][:60], e.get("args", {}).get("stream", -1), e["ts"], e["dur"]) for e in ev if e.get("cat") == "kernel"]
rows.sort(key=lambda r: r[2])
t0 = rows[0][2] if rows else 0
json.dump([(n, s, round(ts - t0, 3), round(d, 3)) for n, s, ts, d in rows], open(out, "w"))
os.remove(tmp)
print("kernels traced:", len(rows))
