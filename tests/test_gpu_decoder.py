"""Decoder step / greedy loop against the fp32 oracle: teacher-forced logits within the bf16 tolerance, tokens identical
except where the reference's top-2 margin is below it; model-ABI entry points (decoder_main / decoder_loop)."""
import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["micro", "tiny"])
def setup(request, pkg):
    arch = request.param
    B = 2
    eng = pkg.Engine(util.model_root(arch), arch, 0, B)
    oracle = util.load_oracle(arch)
    audios = [util.synth_audio("S", 480000, 7), util.synth_audio("N", 200000, 8)]
    mel = eng.logmel(audios)
    eng.encoder(batch=B, return_cross=False)
    with torch.no_grad():
        ck, cv = oracle.encoder(mel)
    yield eng, oracle, ck, cv, B
    eng.close()


def test_teacher_forced_logits(setup):
    eng, oracle, ck, cv, B = setup
    n = 24
    ref = oracle.greedy(ck, cv, max_new_tokens=n, honor_eot=False, keep_logits=True)
    forced = np.array(ref["tokens"], np.int32)
    toks, logits = eng.greedy(B, max_new_tokens=n, honor_eot=False, forced_tokens=forced, keep_logits=True)
    ref_logits = np.stack(ref["logits"])  # [n, B, V]
    err = np.abs(logits[:n] - ref_logits).max()
    print("teacher-forced logits max-abs err %.4f (max |logit| %.2f)" % (err, np.abs(ref_logits).max()))
    assert err <= util.logit_tol(ref_logits)
    margins = np.stack(ref["top2_margin"])
    bad = 0
    for i in range(n):
        for b in range(B):
            if toks[b][i] != ref["tokens"][b][i]:
                assert margins[i][b] < 2 * err + 1e-6, "argmax differs at a step with margin %.4f" % margins[i][b]
                bad += 1
    print("argmax disagreements under the margin rule: %d of %d" % (bad, n * B))


def test_free_running_tokens(setup):
    eng, oracle, ck, cv, B = setup
    n = 32
    ref = oracle.greedy(ck, cv, max_new_tokens=n, honor_eot=False, keep_logits=True)
    toks, _ = eng.greedy(B, max_new_tokens=n, honor_eot=False)
    # a flip needs both logits to move towards each other: margin threshold = 2 x the logit tolerance
    tol = 2 * util.logit_tol(np.stack(ref["logits"]))
    assert util.tokens_agree(toks, ref["tokens"], ref["top2_margin"], tol), (toks, ref["tokens"])
    same = sum(a == b for x, y in zip(toks, ref["tokens"]) for a, b in zip(x, y))
    print("free-running: %d of %d tokens identical to the oracle" % (same, n * B))
    # CUDA-graph path and eager path are the same kernels: identical output
    toks2, _ = eng.greedy(B, max_new_tokens=n, honor_eot=False, keep_logits=True)
    assert toks == toks2


def test_model_abi_steps(setup):
    """decoder_main == 4 SOT steps, decoder_loop == one step; this_self_k/v rows match the oracle's."""
    eng, oracle, ck, cv, B = setup
    sot = eng.sot_sequence("zh")
    assert sot == oracle.sot_sequence("zh")
    logits, k4, v4 = eng.decoder_main(sot, B)
    L, d = oracle.l_dec, oracle.d
    self_k = torch.zeros(L, B, 448, d)
    self_v = torch.zeros(L, B, 448, d)
    mask = torch.ones(448, dtype=torch.int32)
    with torch.no_grad():
        for i, t in enumerate(sot):
            if i > 0:
                mask[i - 1] = 0
            rl, rk, rv = oracle.decoder_step(torch.full((B,), t), self_k, self_v, ck, cv, i, mask)
            self_k[:, :, i], self_v[:, :, i] = rk, rv
            assert np.abs(k4[:, :, i] - rk.numpy()).max() <= 2e-2 * max(1.0, float(rk.abs().max()))
            assert np.abs(v4[:, :, i] - rv.numpy()).max() <= 2e-2 * max(1.0, float(rv.abs().max()))
        assert np.abs(logits - rl.numpy()).max() <= util.logit_tol(rl.numpy())
        nxt = rl.argmax(-1)
        mask[3] = 0
        rl2, rk2, _ = oracle.decoder_step(nxt, self_k, self_v, ck, cv, 4, mask)
    l2, k1, _ = eng.decoder_loop(nxt.numpy().astype(np.int32), 4)
    assert np.abs(l2 - rl2.numpy()).max() <= util.logit_tol(rl2.numpy())
    assert np.abs(k1 - rk2.numpy()).max() <= 2e-2 * max(1.0, float(rk2.abs().max()))


def test_eot_is_honoured(setup):
    eng, oracle, ck, cv, B = setup
    ref = oracle.greedy(ck, cv, max_new_tokens=40, honor_eot=True, keep_logits=True)
    toks, _ = eng.greedy(B, max_new_tokens=40, honor_eot=True)
    eot = int(oracle.cfg["eot"])
    assert all(eot not in t for t in toks)
    assert util.tokens_agree(toks, ref["tokens"], ref["top2_margin"], 2 * util.logit_tol(np.stack(ref["logits"])))


def test_stateless_decoder_step(setup, pkg):
    """The reference decoder's own contract (export_onnx.py:668-670, Whisper.cpp:306-326): tokens, self_k/v, cross_k/v,
    offset and mask go IN, logits and this_self_k/v come OUT.  Every cache is the ORACLE's here, so the step is compared
    with the encoder and the previous steps taken out; the updated cache is read back through the same boundary."""
    eng, oracle, ck, cv, B = setup
    L, d = oracle.l_dec, oracle.d
    sot = oracle.sot_sequence("zh")
    self_k = torch.zeros(L, B, 448, d)
    self_v = torch.zeros(L, B, 448, d)
    mask = torch.ones(448, dtype=torch.int32)
    toks = [torch.full((B,), t) for t in sot] + [torch.tensor([100 + b for b in range(B)]), torch.tensor([2000 + 7 * b for b in range(B)])]
    with torch.no_grad():
        for i, t in enumerate(toks[:-1]):
            if i > 0:
                mask[i - 1] = 0
            _, rk, rv = oracle.decoder_step(t, self_k, self_v, ck, cv, i, mask)
            self_k[:, :, i], self_v[:, :, i] = rk, rv
        off = len(toks) - 1
        mask[off - 1] = 0
        rl, rk, rv = oracle.decoder_step(toks[-1], self_k, self_v, ck, cv, off, mask)
    logits, k1, v1 = eng.decoder_step(toks[-1].numpy(), off, self_k=self_k.numpy(), self_v=self_v.numpy(), cross_k=ck.numpy(),
                                      cross_v=cv.numpy(), mask=mask.numpy())
    err = np.abs(logits - rl.numpy()).max()
    print("stateless step: logits max-abs err %.4f (tol %.4f)" % (err, util.logit_tol(rl.numpy())))
    assert err <= util.logit_tol(rl.numpy())
    assert np.abs(k1 - rk.numpy()).max() <= 2e-2 * max(1.0, float(rk.abs().max()))
    assert np.abs(v1 - rv.numpy()).max() <= 2e-2 * max(1.0, float(rv.abs().max()))
    # the cache that went in comes back (bf16-rounded) with the new row appended at `off`
    sk, sv = eng.get_self_kv(B, off + 1)
    assert np.abs(sk[:, :, :off] - self_k[:, :, :off].numpy()).max() <= 2 ** -8 * float(self_k.abs().max())
    assert np.abs(sk[:, :, off] - k1).max() <= 2 ** -7 * max(1.0, float(np.abs(k1).max()))
    assert np.abs(sv[:, :, off] - v1).max() <= 2 ** -7 * max(1.0, float(np.abs(v1).max()))
    # keeping the resident caches (NULL inputs) and stepping again == the stateful decoder_loop
    nxt = np.array([31 + b for b in range(B)], np.int32)
    l_a, _, _ = eng.decoder_step(nxt, off + 1)
    l_b, _, _ = eng.decoder_loop(nxt, off + 1)
    assert np.array_equal(l_a, l_b)
    # a mask that is not the causal one of Whisper.cpp:253-258 is refused, not silently ignored
    bad = mask.numpy().copy()
    bad[0] = 1
    with pytest.raises(pkg.B200Error):
        eng.decoder_step(nxt, off + 1, mask=bad)


def test_batch_beyond_capacity_is_an_error(setup, pkg):
    """greedy / decoder_main / decoder_loop on more sequences than the resident workspace holds must fail
    cleanly instead of writing out of bounds."""
    eng, oracle, ck, cv, B = setup
    with pytest.raises(pkg.B200Error):
        eng.greedy(B + 7, max_new_tokens=2, honor_eot=False)
    with pytest.raises(pkg.B200Error):
        eng.decoder_loop(np.zeros(B + 7, np.int32), 4)
    with pytest.raises(pkg.B200Error):
        eng.decoder_main(eng.sot_sequence("zh"), B + 7)
    toks, _ = eng.greedy(B, max_new_tokens=2, honor_eot=False)  # the engine is still usable
    assert len(toks) == B
