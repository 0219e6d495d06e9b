"""BASELINE-sized checks through size-independent properties (the fp32 oracle needs ~10 s per Whisper-base chunk, so the
full configs are not compared element-wise): Whisper-base, batch 64 x 30 s (configs[1]).
  * batch invariance: a chunk's tokens do not depend on its position in the batch or on the other chunks;
  * micro-batch / stream / CUDA-graph invariance: two-stream graph decode == single-stream eager decode;
  * determinism run to run; greedy decode length bookkeeping; the first chunks agree with the oracle."""
import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu


def test_base_batch64_properties(pkg):
    arch, B, n_new = "base", 64, 24
    eng = pkg.Engine(util.model_root(arch), arch, 0, B)
    base = [util.synth_audio("NUS"[i % 3], 480000 if i % 5 else 250000 + 1000 * i, 300 + i) for i in range(8)]
    audios = [base[i % 8] for i in range(B)]
    toks, times = eng.transcribe(audios, max_new_tokens=n_new, honor_eot=False)
    assert all(len(t) == n_new for t in toks)
    assert times["decode_steps"] == 4 + n_new
    for i in range(B):
        assert toks[i] == toks[i % 8], "chunk %d differs from its copy at position %d" % (i, i % 8)
    toks2, _ = eng.transcribe(audios, max_new_tokens=n_new, honor_eot=False)
    assert toks2 == toks
    # permuted batch
    perm = np.random.default_rng(0).permutation(B)
    toks3, _ = eng.transcribe([audios[j] for j in perm], max_new_tokens=n_new, honor_eot=False)
    assert [toks3[k] for k in range(B)] == [toks[j] for j in perm]
    # eager (no graph) path with logits kept == graph path
    eng.logmel(audios)
    eng.encoder(batch=B, return_cross=False)
    toks4, _ = eng.greedy(B, max_new_tokens=n_new, honor_eot=False, keep_logits=True)
    assert toks4 == toks
    # oracle on the first two chunks
    oracle = util.load_oracle(arch)
    mel = eng.logmel(audios[:2])
    with torch.no_grad():
        ref = oracle.transcribe_tokens(mel, max_new_tokens=n_new, honor_eot=False, keep_logits=True)
    tol = 2 * util.logit_tol(np.stack(ref["logits"]))
    assert util.tokens_agree(toks[:2], ref["tokens"], ref["top2_margin"], tol)
    eng.close()
