"""K1 log-mel kernel against the reference's own frontend (oracle/_ref) on seeded audio; tolerance 1e-4 absolute."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engines(pkg):
    e = {80: pkg.Engine(util.model_root("micro"), "micro", 0, 4)}
    yield e
    for v in e.values():
        v.close()


@pytest.mark.parametrize("dist", ["N", "U", "S", "T"])
def test_mel_30s(engines, dist):
    a = util.synth_audio(dist, 480000, seed=100 + ord(dist))
    got = engines[80].logmel([a])
    ref = util.reference_mel([a], 80)
    err = np.abs(got - ref).max()
    print("dist %s max-abs %.3e" % (dist, err))
    assert err <= util.MEL_TOL


def test_mel_ragged_batch_and_short_clip(engines):
    # demo.wav shape (67263 samples: zero fill after normalisation), a 201-sample minimum clip, > 30 s (max over all frames)
    audios = [util.synth_audio("S", 67263, 1), util.synth_audio("N", 201, 2), util.synth_audio("U", 480000, 3),
              util.synth_audio("S", 500000, 4)]
    got = engines[80].logmel(audios)
    ref = util.reference_mel(audios, 80)
    for i in range(len(audios)):
        assert np.abs(got[i] - ref[i]).max() <= util.MEL_TOL, "utterance %d" % i
    n_frames = 1 + 67263 // 160
    assert np.all(got[0][:, n_frames:] == 0.0)




def test_mel_golden(engines):
    import os
    g = np.load(os.path.join(util.ROOT, "tests", "golden", "mel_golden.npz"))
    a = util.synth_audio("S", 67263, 1)
    got = engines[80].logmel([a])[0]
    assert np.abs(got[:, :430] - g["short_S_80"]).max() <= util.MEL_TOL
