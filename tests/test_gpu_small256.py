"""The headline configuration (BASELINE.json configs[2]: Whisper-small, 256 x 30 s chunks on one GPU) through
size-independent properties, on the code path bench.py times: two micro-batches of 128 sequences, the streaming
cross-attention kernel, thin decoder GEMMs, eight decoder steps per CUDA graph.
  * batch / micro-batch invariance: a chunk's tokens do not depend on its slot (copies land in both micro-batches);
  * determinism run to run (the streaming kernel claims its work items dynamically);
  * multi-step graphs == eager single-step decode;
  * the first chunk agrees with the fp32 oracle under the top-2-margin rule."""
import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu


def test_small_batch256_properties(pkg):
    arch, B, n_new = "small", 256, 20  # 24 decoder steps = three 8-step graphs
    eng = pkg.Engine(util.model_root(arch), arch, 0, B)
    base = [util.synth_audio("NUS"[i % 3], 480000 if i % 4 else 200000 + 3000 * i, 500 + i) for i in range(8)]
    audios = [base[i % 8] for i in range(B)]
    toks, times = eng.transcribe(audios, max_new_tokens=n_new, honor_eot=False)
    assert all(len(t) == n_new for t in toks)
    assert times["decode_steps"] == 4 + n_new
    for i in range(B):
        assert toks[i] == toks[i % 8], "chunk %d differs from its copy at position %d" % (i, i % 8)
    toks2, _ = eng.transcribe(audios, max_new_tokens=n_new, honor_eot=False)
    assert toks2 == toks
    # 21 new tokens = 25 steps: three 8-step graphs + one single-step graph give the same prefix
    toks3, _ = eng.transcribe(audios, max_new_tokens=n_new + 1, honor_eot=False)
    assert [t[:n_new] for t in toks3] == toks
    # eager (no graph, one step per enqueue) path
    eng.logmel(audios)
    eng.encoder(batch=B, return_cross=False)
    toks4, _ = eng.greedy(B, max_new_tokens=8, honor_eot=False, keep_logits=True)  # keeping logits forces the eager path
    assert toks4 == [t[:8] for t in toks]
    # oracle on the first chunk
    oracle = util.load_oracle(arch)
    mel = eng.logmel(audios[:1])
    with torch.no_grad():
        ref = oracle.transcribe_tokens(mel, max_new_tokens=8, honor_eot=False, keep_logits=True)
    tol = 2 * util.logit_tol(np.stack(ref["logits"]))
    assert util.tokens_agree([toks[0][:8]], ref["tokens"], ref["top2_margin"], tol)
    eng.close()

