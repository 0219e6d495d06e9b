#!/usr/bin/env python3
"""BASELINE.json configs[4]: 1 h of synthetic 16 kHz audio cut into 30 s windows, Whisper-small, through the reference-shaped
C API (AX_WHISPER_RunPCMLong) on every visible GPU of the box (one process, one engine + host thread per GPU, B200W_DEVICES=all).
Prints one JSON line: wall time, RTF (wall / audio duration) including the mel frontend and the host<->device copies."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import __graft_entry__ as g
import util

arch = sys.argv[1] if len(sys.argv) > 1 else "small"
hours = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
n_gpu = torch.cuda.device_count()
os.environ["B200W_DEVICES"] = "all"
os.environ.setdefault("B200W_MAX_BATCH", str(max(1, int(hours * 120 / max(1, n_gpu)) + 1)))
pkg = g.load_package()
w = pkg.Whisper(arch, util.model_root(arch), "zh")
base = np.concatenate([util.synth_audio("NUS"[i % 3], 480000, 900 + i) for i in range(6)])  # 3 min of distinct audio
audio = np.tile(base, int(np.ceil(hours * 20)))[: int(hours * 3600 * 16000)]
w.run_long(audio[: 480000 * max(2, 2 * n_gpu)])  # warm: graphs, workspaces
times = []
for _ in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    text = w.run_long(audio)
    times.append(time.perf_counter() - t0)
wall = min(times)
print(json.dumps({"workload": "long-form %s, %.2f h synthetic audio, %d x 30 s windows, greedy until EOT / 448-token context (random-init weights never emit EOT: 444 tokens per window)" % (arch, hours, int(np.ceil(len(audio) / 480000))),
                  "n_gpus": n_gpu, "wall_s": wall, "rtf": wall / (len(audio) / 16000.0), "audio_s_per_s": (len(audio) / 16000.0) / wall,
                  "text_chars": len(text)}))
w.close()
