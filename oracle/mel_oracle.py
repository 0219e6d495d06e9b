"""CPU restatement (numpy) of the reference's log-mel frontend.  TEST INFRASTRUCTURE ONLY.

Follows SURVEY.md App. A.1 step by step:
  reflect pad 200          /root/reference/cpp/src/librosa/librosa.h:46-57
  periodic Hann, 400/160   librosa.h:79-96 (frames = 1 + n // 160)
  |X|^2                    librosa.h:98-100
  Slaney mel bank (float)  librosa.h:102-144, int fmin/fmax (:112, :146-149)
  mel = M @ P^T            librosa.h:153
  log10 / max / clamp / (x+4)/4 / crop-or-zero-fill to 3000      /root/reference/cpp/src/Whisper.cpp:151-184
It is an independent re-derivation in float64/float32 numpy, NOT bit-compatible with the reference's fp32 kissfft:
it agrees with oracle/_ref (the compiled reference) to ~1e-5 on broadband audio and up to ~1e-4 on tonal audio
(SURVEY.md App. C.3).  Pinned against oracle/_ref outputs through tests/golden/mel_*.npz (tools/make_golden.py).
The GPU parity tests use oracle/_ref itself whenever it is present and fall back to this port otherwise.
"""
import numpy as np

N_FFT, HOP, N_OUT = 400, 160, 3000


def mel_bank(n_mels, sr=16000, n_fft=N_FFT, fmin=0, fmax=8000):
    """float32 arithmetic in the order of librosa.h:102-144."""
    f32 = np.float32
    n_f = n_fft // 2 + 1
    fft_freqs = (np.arange(n_f, dtype=f32) * f32(sr)) / f32(n_fft)
    f_min, f_sp = f32(0.0), f32(200.0) / f32(3.0)
    min_log_hz = f32(1000.0)
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = f32(np.log(f32(6.4))) / f32(27.0)

    def hz_to_mel(hz):  # int argument (librosa.h:112)
        hz = int(hz)
        mel = (f32(hz) - f_min) / f_sp
        if hz >= min_log_hz:
            mel = min_log_mel + f32(np.log(f32(hz) / min_log_hz)) / logstep
        return f32(mel)

    lo, hi = hz_to_mel(fmin), hz_to_mel(fmax)
    step = (hi - lo) / f32(n_mels + 1)
    mels = (lo + np.arange(n_mels + 2, dtype=f32) * step).astype(f32)
    mels[-1] = hi
    mel_f = np.where(mels > min_log_mel, np.exp((mels - min_log_mel) * logstep).astype(f32) * min_log_hz, mels * f_sp + f_min).astype(f32)
    fdiff = (mel_f[1:] - mel_f[:-1]).astype(f32)
    ramps = (mel_f[:, None] - fft_freqs[None, :]).astype(f32)
    lower = (-ramps[:n_mels] / fdiff[:n_mels, None]).astype(f32)
    upper = (ramps[2:] / fdiff[1:, None]).astype(f32)
    w = np.maximum(np.minimum(lower, upper), f32(0))
    enorm = (2.0 / (mel_f[2:].astype(np.float64) - mel_f[:n_mels].astype(np.float64))).astype(f32)
    return (w * enorm[:, None]).astype(f32)


def log_mel(x, n_mels):
    x = np.asarray(x, np.float32)
    n = len(x)
    assert n >= 201
    xp = np.concatenate([x[1:201][::-1], x, x[n - 201:n - 1][::-1]])  # x[200-i], x, x[n-2-k]
    n_frames = 1 + (len(xp) - N_FFT) // HOP
    window = (0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(N_FFT) / N_FFT))).astype(np.float32)
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(n_frames)[:, None]
    frames = (xp[idx] * window[None, :]).astype(np.float32)
    X = np.fft.rfft(frames.astype(np.float64), axis=1)
    P = (X.real ** 2 + X.imag ** 2).astype(np.float32)
    mel = (mel_bank(n_mels).astype(np.float64) @ P.T.astype(np.float64)).astype(np.float32)  # [n_mels, n_frames]
    L = np.log10(np.maximum(mel, np.float32(1e-10))).astype(np.float32)
    mmax = L.max()                                             # every row, every frame (Whisper.cpp:157-167)
    floor = np.float32(np.float64(mmax) - 8.0)
    out = ((np.maximum(L, floor).astype(np.float64) + 4.0) / 4.0).astype(np.float32)
    res = np.zeros((n_mels, N_OUT), np.float32)                # zero fill AFTER normalisation (:172)
    keep = min(n_frames, N_OUT)
    res[:, :keep] = out[:, :keep]
    return res
