"""CPU checks of the log-mel oracles: the compiled reference frontend (oracle/_ref, when present) must reproduce the
committed golden vectors bit for bit, and the numpy restatement must agree with them within the parity gate."""
import os

import numpy as np
import pytest

import mel_oracle
import util

GOLD = np.load(os.path.join(util.ROOT, "tests", "golden", "mel_golden.npz"))
STRIDE = 25


def test_numpy_port_matches_reference_golden_short_clip():
    a = util.synth_audio("S", 67263, 1)  # demo.wav shape: 421 frames, zero fill after normalisation
    got = mel_oracle.log_mel(a, 80)
    assert got.shape == (80, 3000)
    assert np.abs(got[:, :430] - GOLD["short_S_80"]).max() <= util.MEL_TOL
    assert np.all(got[:, 1 + 67263 // 160:] == 0.0)


@pytest.mark.parametrize("dist", ["N", "U", "S"])
@pytest.mark.parametrize("n_mels", [80, 128])
def test_numpy_port_matches_reference_golden_30s(dist, n_mels):
    a = util.synth_audio(dist, 480000, seed=100 + ord(dist))
    got = mel_oracle.log_mel(a, n_mels)[:, ::STRIDE]
    assert np.abs(got - GOLD["full_%s_%d" % (dist, n_mels)]).max() <= util.MEL_TOL


@pytest.mark.skipif(util.mel_ref_lib() is None, reason="oracle/_ref not built (needs /root/reference)")
def test_compiled_reference_reproduces_golden_bitwise():
    a = util.synth_audio("S", 67263, 1)
    assert np.array_equal(util.reference_mel([a], 80)[0][:, :430], GOLD["short_S_80"])
    a = util.synth_audio("U", 480000, seed=100 + ord("U"))
    assert np.array_equal(util.reference_mel([a], 128)[0][:, ::STRIDE], GOLD["full_U_128"])


def test_mel_bank_restatement_is_close_to_reference_tables(pkg):
    # the product's committed constant tables (generated from the reference build) vs the float32 restatement
    for n_mels in (80, 128):
        bank, win = pkg.mel_tables(n_mels)
        mine = mel_oracle.mel_bank(n_mels)
        assert bank.shape == mine.shape
        assert np.abs(bank - mine).max() <= 2e-6
        assert np.count_nonzero(bank) > 2 * n_mels
    w = 0.5 * (1 - np.cos(2 * np.pi * np.arange(400) / 400))
    assert np.abs(win - w).max() <= 1e-6  # Eigen evaluates the window in float with a vectorised cos


@pytest.mark.skipif(util.mel_ref_lib() is None, reason="oracle/_ref not built (needs /root/reference)")
def test_committed_tables_equal_reference_build(pkg):
    import ctypes
    lib = util.mel_ref_lib()
    fp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    win = np.zeros(400, np.float32)
    lib.melref_window(fp(win))
    for n_mels in (80, 128):
        ref = np.zeros((n_mels, 201), np.float32)
        lib.melref_bank(ctypes.c_int(n_mels), fp(ref))
        bank, w = pkg.mel_tables(n_mels)
        assert np.array_equal(bank, ref)
        assert np.array_equal(w, win)


def test_edge_cases():
    with pytest.raises(AssertionError):
        mel_oracle.log_mel(np.zeros(200, np.float32), 80)  # reflect padding needs 201 samples
    z = mel_oracle.log_mel(np.zeros(16000, np.float32), 80)  # silence: log10(1e-10) = -10 everywhere -> (-10 + 4) / 4
    assert np.allclose(z[:, :101], -1.5) and np.all(z[:, 101:] == 0)
