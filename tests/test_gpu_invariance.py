"""A sequence's tokens do not depend on the batch it is decoded in (VERDICT r01 "output depends on the shard size").

The same 16 chunks are transcribed alone (B = 16), scattered inside a batch of 64 and inside a batch of 256.  These batch
sizes take every kernel variant the engine has: cluster-split cross attention with different split factors vs the streaming
kernel, one vs two decoder micro-batches, one vs two encoder sub-batches, several GEMM tile counts.  All variants share one
summation order (decode_ops.cu), so the token streams must be IDENTICAL -- no margin rule here.  The 1-GPU vs 2-GPU split of
the same batch is tests/test_gpu_multi.py (needs two GPUs)."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("arch,sizes", [("base", (16, 64, 256)), ("small", (16, 40, 256)), ("tiny", (1, 16, 200))])
def test_tokens_do_not_depend_on_batch_size(pkg, arch, sizes):
    n_new = 40
    eng = pkg.Engine(util.model_root(arch), arch, 0, max(sizes))
    n_probe = min(16, min(sizes))
    probe = [util.synth_audio("NUS"[i % 3], 480000 if i % 3 else 260000 + 9000 * i, 4000 + i) for i in range(n_probe)]
    filler = [util.synth_audio("NUS"[(i + 1) % 3], 480000 if i % 2 else 350000, 5000 + i) for i in range(8)]
    ref = None
    for B in sizes:
        slots = [int(x) for x in np.linspace(0, B - 1, n_probe).astype(int)] if B > n_probe else list(range(n_probe))
        assert len(set(slots)) == n_probe
        audios = [filler[i % 8] for i in range(B)]
        for j, s in enumerate(slots):
            audios[s] = probe[j]
        toks, _ = eng.transcribe(audios, max_new_tokens=n_new, honor_eot=False)
        got = [toks[s] for s in slots]
        assert all(len(t) == n_new for t in got)
        if ref is None:
            ref = got
            print("%s: %d distinct tokens in the %d probe sequences" % (arch, len({t for x in got for t in x}), n_probe))
        else:
            bad = [(j, next(i for i in range(n_new) if got[j][i] != ref[j][i])) for j in range(n_probe) if got[j] != ref[j]]
            assert not bad, "%s: probe sequences differ between B=%d and B=%d at (sequence, first step): %s" % (arch, sizes[0], B, bad)
    eng.close()
