#!/usr/bin/env python3
"""Are the tokens independent of timing?  Runs the headline workload alone, then again while ANOTHER PROCESS keeps the same GPU busy
(contexts are time-sliced: kernels get preempted at arbitrary points), and compares the token ids.  Env switches passed on the command
line (e.g. B200W_NO_PDL=1) apply to the engine; usage: debug_disturb.py ARCH BATCH [KEY=VALUE ...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

if len(sys.argv) > 1 and sys.argv[1] == "--spin":
    import torch

    a = torch.randn(4096, 4096, device="cuda")
    mode = sys.argv[2]
    t_end = time.time() + float(sys.argv[3])
    while time.time() < t_end:
        if mode == "sleep":
            torch.cuda._sleep(200000000)  # one long-running spinning kernel (what an NCCL barrier looks like)
        else:
            for _ in range(20):
                a = (a @ a).clamp_(-1, 1)
        torch.cuda.synchronize()
    sys.exit(0)

import numpy as np

import __graft_entry__ as g
import bench
import util

arch, B = sys.argv[1], int(sys.argv[2])
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    os.environ[k] = v
pkg = g.load_package()
eng = pkg.Engine(util.model_root(arch), arch, 0, B)
pcm = np.stack([bench.synth_chunk(i) for i in range(B)])
host = os.environ.get("DISTURB_HOST_PATH") == "1"  # the C-API path: pageable host PCM, pipelined H2D copies
if host:
    run = lambda: eng.transcribe(pcm, max_new_tokens=224, honor_eot=False)
else:
    eng.upload_pcm(pcm)
    run = lambda: eng.transcribe_resident(B, max_new_tokens=224, honor_eot=False)
base, _ = run()
for rep in range(2):
    t, _ = run()
    assert t == base, "not deterministic even when alone"
for mode in ("sleep", "matmul"):
    child = subprocess.Popen([sys.executable, os.path.abspath(__file__), "--spin", mode, "25"])
    time.sleep(8)  # let the child create its context and start spinning
    bad_runs = 0
    for rep in range(4):
        t0 = time.time()
        t, _ = run()
        bad = [(i, next(k for k in range(224) if t[i][k] != base[i][k])) for i in range(B) if t[i] != base[i]]
        print("%s %s disturbed run %d: %.2f s, mismatching sequences %d %s" % (" ".join(sys.argv[1:]), mode, rep, time.time() - t0, len(bad), bad[:6]), flush=True)
        bad_runs += bool(bad)
    child.wait()
eng.close()
