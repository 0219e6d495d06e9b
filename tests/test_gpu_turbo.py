"""Whisper-turbo shape (large-v3-turbo: 128 mel bins, d = 1280, 20 heads, 32 encoder / 4 decoder layers, 51866-token vocabulary,
BASELINE.json configs[3]) through the whole path.  The fp32 oracle needs ~1 minute per chunk at this size, so only the
128-bin log-mel is compared element-wise; the rest is checked through properties (determinism, batch invariance,
graph == eager) plus one teacher-forced oracle comparison on a single short chunk."""
import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu


def test_turbo_end_to_end(pkg):
    arch = "turbo"
    B = 4
    eng = pkg.Engine(util.model_root(arch), arch, 0, B)
    assert (eng.dims.n_mels, eng.dims.d_model, eng.dims.n_head, eng.dims.n_audio_layer, eng.dims.n_text_layer, eng.dims.n_vocab) == (128, 1280, 20, 32, 4, 51866)
    assert eng.sot_sequence("zh") == [50258, 50260, 50360, 50364]  # SURVEY.md App. A.5 (100-language tokenizer)
    audios = [util.synth_audio("S", 480000, 61), util.synth_audio("N", 300000, 62), util.synth_audio("U", 480000, 63), util.synth_audio("S", 480000, 61)]
    mel = eng.logmel(audios)
    ref = util.reference_mel(audios, 128)
    assert np.abs(mel - ref).max() <= util.MEL_TOL
    toks, _ = eng.transcribe(audios, max_new_tokens=16, honor_eot=False)
    assert toks[0] == toks[3] and all(len(t) == 16 for t in toks)
    toks2, _ = eng.transcribe(audios, max_new_tokens=16, honor_eot=False)
    assert toks2 == toks
    eng.logmel(audios)
    eng.encoder(batch=B, return_cross=False)
    toks3, _ = eng.greedy(B, max_new_tokens=16, honor_eot=False, keep_logits=True)  # eager path
    assert toks3 == toks
    # oracle (fp32, CPU) on the first chunk only
    oracle = util.load_oracle(arch)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    with torch.no_grad():
        r = oracle.transcribe_tokens(mel[:1], max_new_tokens=8, honor_eot=False, keep_logits=True)
    tol = 2 * util.logit_tol(np.stack(r["logits"]))
    assert util.tokens_agree([toks[0][:8]], r["tokens"], r["top2_margin"], tol)
    eng.close()
