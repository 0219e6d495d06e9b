"""Encoder graph (conv stem, blocks, ln_post, cross K/V projections) against the fp32 oracle, bf16 tolerance."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def _check(got, ref, name):
    rel = np.abs(got - ref).max() / np.abs(ref).max()
    cos = util.cosine(got, ref)
    print("%s: max-abs-err/max-abs-ref %.4f cosine %.6f" % (name, rel, cos))
    assert rel <= util.ENC_REL_TOL, name
    assert cos >= util.ENC_COS_TOL, name


@pytest.mark.parametrize("arch,B", [("micro", 2), ("tiny", 1), ("tiny", 3)])
def test_encoder_cross_kv(pkg, arch, B):
    eng = pkg.Engine(util.model_root(arch), arch, 0, B)
    oracle = util.load_oracle(arch)
    audios = [util.synth_audio("NUS"[i % 3], 480000 if i != 1 else 300000, seed=40 + i) for i in range(B)]
    mel = eng.logmel(audios)
    ck, cv = eng.encoder(batch=B)  # resident mel
    import torch
    with torch.no_grad():
        rk, rv = oracle.encoder(mel)
    _check(ck, rk.numpy(), "cross_k")
    _check(cv, rv.numpy(), "cross_v")
    # entering with a caller-supplied mel tensor (model ABI) gives the same result as the fused path
    ck2, cv2 = eng.encoder(mel=mel)
    assert np.array_equal(ck, ck2) and np.array_equal(cv, cv2)
    eng.close()
