"""Shared test helpers: seeded synthetic audio, model directories, oracle access, comparison rules."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import make_model  # noqa: E402

MODEL_CACHE = os.environ.get("B200W_MODEL_CACHE", "/tmp/b200w_models")
MEL_TOL = 1e-4          # north_star: mel within 1e-4 absolute
LOGIT_REL_TOL = 1.5e-2  # bf16 tolerance on logits: max-abs error <= 1.5 % of the largest |logit| of the compared block
                        # (weights, activations and KV caches are bf16 = 8 mantissa bits; the max is over ~10^6 logits)
ENC_REL_TOL = 2e-2      # max-abs error of cross K/V relative to max-abs of the reference tensor
ENC_COS_TOL = 0.999


def synth_audio(dist, n, seed):
    """N = 0.1*N(0,1) clipped, U = U(-1,1), S = 5 AM sines 80-4000 Hz + 0.01*N(0,1)."""
    rng = np.random.default_rng(seed)
    if dist == "N":
        return np.clip(0.1 * rng.standard_normal(n), -1, 1).astype(np.float32)
    if dist == "U":
        return rng.uniform(-1, 1, n).astype(np.float32)
    if dist == "S":
        t = np.arange(n) / 16000.0
        x = np.zeros(n)
        for _ in range(5):
            f = rng.uniform(80, 4000)
            am = rng.uniform(2, 8)
            x += rng.uniform(0.05, 0.2) * np.sin(2 * np.pi * f * t + rng.uniform(0, 6.28)) * (0.6 + 0.4 * np.sin(2 * np.pi * am * t))
        x += 0.01 * rng.standard_normal(n)
        return np.clip(x, -1, 1).astype(np.float32)
    if dist == "T":  # adversarial: pure tones, no noise floor
        t = np.arange(n) / 16000.0
        return (0.5 * np.sin(2 * np.pi * 440.0 * t) + 0.3 * np.sin(2 * np.pi * 1234.5 * t)).astype(np.float32)
    raise ValueError(dist)


def model_root(arch):
    """Builds (once per box) the seeded random-init model directory for `arch`; returns the model root."""
    d = os.path.join(MODEL_CACHE, arch)
    marker = os.path.join(d, ".complete")
    if not os.path.exists(marker):
        make_model.build_model_dir(MODEL_CACHE, arch)
        open(marker, "w").write("ok")
    return MODEL_CACHE


_oracles = {}


def load_oracle(arch):
    import whisper_oracle

    if arch not in _oracles:
        W, cfg = make_model.load_model_dir(model_root(arch), arch)
        _oracles[arch] = whisper_oracle.Oracle(W, cfg)
    return _oracles[arch]


_melref = None


def mel_ref_lib():
    """oracle/_ref/libmel_ref.so (the reference's own frontend) or None if it was not built."""
    global _melref
    if _melref is None:
        path = os.path.join(ROOT, "oracle", "_ref", "libmel_ref.so")
        _melref = ctypes.CDLL(path) if os.path.exists(path) else False
    return _melref or None


def reference_mel(audios, n_mels):
    """[B, n_mels, 3000] from the reference frontend when built, else from the numpy restatement."""
    lib = mel_ref_lib()
    out = np.zeros((len(audios), n_mels, 3000), np.float32)
    if lib is not None:
        for i, a in enumerate(audios):
            a = np.ascontiguousarray(a, np.float32)
            rc = lib.melref_preprocess(a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), len(a), n_mels,
                                       out[i].ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
            assert rc > 0
        return out
    import mel_oracle

    for i, a in enumerate(audios):
        out[i] = mel_oracle.log_mel(a, n_mels)
    return out


def logit_tol(ref_logits):
    return LOGIT_REL_TOL * float(np.abs(ref_logits).max())


def token_report(got, exp, margins, tol):
    """Per sequence: how many positions were compared before the first divergence, where it diverged and the oracle's
    top-2 margin there.  A sequence is ok when it is identical to the oracle's (same length), or when its first divergence
    sits on a step whose reference top-2 margin is below `tol` (after such a step free-running sequences may legitimately
    differ; a shorter or longer output without such a step is a failure)."""
    rep = []
    for b in range(len(exp)):
        n = min(len(got[b]), len(exp[b]))
        div = next((i for i in range(n) if got[b][i] != exp[b][i]), None)
        if div is None and len(got[b]) != len(exp[b]):
            div = n  # one side stopped early: a divergence at the first missing position
        margin = float(margins[div][b]) if div is not None and div < len(margins) else None
        ok = div is None or (margin is not None and margin < tol)
        rep.append(dict(seq=b, len_got=len(got[b]), len_ref=len(exp[b]), compared=n if div is None else div, first_divergence=div,
                        margin_at_divergence=margin, ok=bool(ok)))
    return rep


def tokens_agree(got, exp, margins, tol, min_compared=1, verbose=True):
    """Greedy sequences must be identical up to the first step whose reference top-2 margin is below `tol`.  Prints what was
    actually compared; fails when a sequence diverges at a confident step, when lengths differ without such a step, or when
    fewer than `min_compared` positions were compared in total (a vacuous pass)."""
    assert len(got) == len(exp), "sequence count differs: %d vs %d" % (len(got), len(exp))
    rep = token_report(got, exp, margins, tol)
    total = sum(r["compared"] for r in rep)
    if verbose:
        for r in rep:
            print("  tokens seq %(seq)d: compared %(compared)d of %(len_ref)d, first divergence %(first_divergence)s, margin there %(margin_at_divergence)s, ok %(ok)s" % r)
        print("  tokens: %d positions compared in total (tol %.4f)" % (total, tol))
    return all(r["ok"] for r in rep) and total >= min_compared


def cosine(a, b):
    a = a.ravel().astype(np.float64)
    b = b.ravel().astype(np.float64)
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))
