#!/usr/bin/env python3
"""Debug helper (2 GPUs): one-device handles on GPU 0 and GPU 1 vs the B200W_DEVICES=0,1 handle on the same chunks."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import __graft_entry__ as g
import bench
import util

arch = sys.argv[1] if len(sys.argv) > 1 else "small"
per = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n_new = 224
pkg = g.load_package()
audios = [bench.synth_chunk(i) for i in range(2 * per)]
os.environ["B200W_MAX_BATCH"] = str(per)


def run(env, chunks):
    for k in ("B200W_DEVICE", "B200W_DEVICES"):
        os.environ.pop(k, None)
    os.environ.update(env)
    w = pkg.Whisper(arch, util.model_root(arch), "zh")
    out = w.run_tokens(chunks, max_new_tokens=n_new, honor_eot=False)
    out2 = w.run_tokens(chunks, max_new_tokens=n_new, honor_eot=False)
    w.close()
    assert out == out2, "not deterministic: %s" % env
    return out


def diff(a, b, name):
    bad = [(i, next(k for k in range(len(a[i])) if a[i][k] != b[i][k])) for i in range(len(a)) if a[i] != b[i]]
    print(name, "mismatching sequences:", len(bad), bad[:12], flush=True)


d0 = run({"B200W_DEVICE": "0"}, audios[:per]) + run({"B200W_DEVICE": "0"}, audios[per:])
d1 = run({"B200W_DEVICE": "1"}, audios[:per]) + run({"B200W_DEVICE": "1"}, audios[per:])
diff(d0, d1, "gpu0 vs gpu1 (separate one-device handles)")
both = run({"B200W_DEVICES": "0,1"}, audios)
diff(d0, both, "gpu0 handle vs two-device handle")
os.environ["B200W_MAX_BATCH"] = str(2 * per)
one = run({"B200W_DEVICE": "0"}, audios)
diff(d0, one, "gpu0 two passes of %d vs one pass of %d" % (per, 2 * per))
