"""The torch restatement of the encoder / decoder graphs reproduces its committed golden outputs (tools/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

import mel_oracle
import util


@pytest.mark.parametrize("arch", ["micro", "tiny"])
def test_oracle_reproduces_golden(arch):
    torch.set_num_threads(8)
    g = np.load(os.path.join(util.ROOT, "tests", "golden", "oracle_%s.npz" % arch))
    o = util.load_oracle(arch)
    audios = [util.synth_audio("S", 480000, 7), util.synth_audio("N", 200000, 8)]
    mel = util.reference_mel(audios, o.n_mels)  # reference frontend if built, numpy port otherwise (<= 1e-4 apart)
    with torch.no_grad():
        ck, cv = o.encoder(mel)
        r = o.greedy(ck, cv, max_new_tokens=24, honor_eot=False, keep_logits=True)
    assert np.abs(ck.numpy()[:, :, ::50, ::4] - g["cross_k"].astype(np.float32)).max() <= 2e-2
    assert np.abs(cv.numpy()[:, :, ::50, ::4] - g["cross_v"].astype(np.float32)).max() <= 2e-2
    logits = np.stack(r["logits"])
    top_val = np.take_along_axis(logits, g["top_idx"].astype(np.int64), -1)
    assert np.abs(top_val - g["top_val"]).max() <= 5e-3
    # tokens identical wherever the golden margin is not a near-tie
    toks = np.array(r["tokens"])
    for b in range(toks.shape[0]):
        for i in range(toks.shape[1]):
            if toks[b, i] != g["tokens"][b, i]:
                assert g["margins"][i, b] < 1e-2
                break


def test_greedy_loop_semantics():
    """Loop bookkeeping of Whisper::run (Whisper.cpp:214-222): 4 SOT steps, at most 444 generated tokens, EOT stops."""
    o = util.load_oracle("micro")
    assert o.sot_sequence("zh") == [50258, 50260, 50359, 50363]
    assert o.sot_sequence("en")[1] == 50259
    assert o.sot_sequence("xx") == o.sot_sequence("zh")  # unknown language falls back to zh (Whisper.cpp:244-248)
    torch.manual_seed(0)
    ck = torch.randn(o.l_dec, 1, 1500, o.d) * 0.5
    cv = torch.randn(o.l_dec, 1, 1500, o.d) * 0.5
    with torch.no_grad():
        r = o.greedy(ck, cv, max_new_tokens=5, honor_eot=False)
        assert len(r["tokens"][0]) == 5
        forced = np.array([[7, 8, 9]])
        r2 = o.greedy(ck, cv, max_new_tokens=5, honor_eot=False, forced_tokens=forced)
    assert r2["tokens"][0][0] == r["tokens"][0][0]  # first prediction does not depend on what is fed back afterwards
