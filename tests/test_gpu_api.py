"""The reference-shaped outer boundary (AX_WHISPER_Init / RunPCM / RunFile / Uninit + batched extensions) on the GPU:
same conventions as /root/reference/cpp/src/api/ax_whisper_api.cpp:48-163, results consistent with the model-level ABI."""
import base64
import ctypes
import os
import struct
import subprocess

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def _write_wav(path, data, sr=16000):
    raw = (np.clip(data, -1, 1) * 32767).astype("<i2").tobytes()
    n_ch = 1 if data.ndim == 1 else data.shape[1]
    hdr = b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, n_ch, sr, sr * n_ch * 2, n_ch * 2, 16)
    open(path, "wb").write(hdr + b"data" + struct.pack("<I", len(raw)) + raw)


@pytest.fixture(scope="module")
def whisper(pkg):
    w = pkg.Whisper("micro", util.model_root("micro"), "zh")
    yield w
    w.close()


def _detok(tokens):
    return "".join(" t%d" % t for t in tokens if t < 50257)


def test_run_pcm_matches_token_api(pkg, whisper):
    a = util.synth_audio("S", 67263, 1)  # demo.wav shape
    text = whisper.run(a)
    toks = whisper.run_tokens([a])[0]
    assert text == _detok(toks)           # synthetic token table: id i decodes to " t<i>"; specials are skipped
    assert len(toks) <= 444               # Whisper.cpp:219: offset < n_text_ctx
    # the model-level ABI gives the same tokens
    eng = pkg.Engine(util.model_root("micro"), "micro", 0, 1)
    toks2, _ = eng.transcribe([a])
    eng.close()
    assert toks == toks2[0]


def test_run_file_and_stereo_mix(whisper, tmp_path):
    a = util.synth_audio("S", 48000, 3)
    b = util.synth_audio("N", 48000, 4)
    _write_wav(str(tmp_path / "mono.wav"), a)
    _write_wav(str(tmp_path / "stereo.wav"), np.stack([a, b], 1))
    q = lambda x: (np.clip(x, -1, 1) * 32767).astype(np.int16).astype(np.float32) / 32768.0  # int16 round trip of the WAV
    assert whisper.run(str(tmp_path / "mono.wav")) == whisper.run(q(a))
    assert whisper.run(str(tmp_path / "stereo.wav")) == whisper.run((q(a) + q(b)) / 2)  # ax_whisper_api.cpp:109-113


def test_batch_equals_single_and_is_deterministic(whisper):
    audios = [util.synth_audio("NUS"[i % 3], 40000 + 7000 * i, 20 + i) for i in range(5)]
    batch = whisper.run_tokens(audios, max_new_tokens=24, honor_eot=False)
    again = whisper.run_tokens(audios, max_new_tokens=24, honor_eot=False)
    assert batch == again
    for i, a in enumerate(audios):
        assert whisper.run_tokens([a], max_new_tokens=24, honor_eot=False)[0] == batch[i]


def test_long_form_windows(whisper):
    """configs[4] shape: audio longer than 30 s is cut into independent 30 s windows; the result is the concatenation of the
    per-window transcriptions, whatever the batch size used to process them."""
    a = np.concatenate([util.synth_audio("S", 480000, 31), util.synth_audio("N", 480000, 32), util.synth_audio("U", 123456, 33)])
    text = whisper.run_long(a)
    parts = [whisper.run(a[i:i + 480000]) for i in range(0, len(a), 480000)]
    assert text == "".join(parts)
    assert whisper.run_long(a, window_batch=2) == text



def test_error_conventions(pkg, whisper, tmp_path):
    lib = pkg.load_library()
    res = ctypes.c_void_p(1)
    short = np.zeros(100, np.float32)
    assert lib.AX_WHISPER_RunPCM(whisper.h, short.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 100, ctypes.byref(res)) == -1
    assert res.value is None              # *result = nullptr before work (ax_whisper_api.cpp:149)
    assert lib.AX_WHISPER_RunFile(whisper.h, b"/nonexistent.wav", ctypes.byref(res)) == -1
    assert lib.AX_WHISPER_Init(b"nope", str(tmp_path).encode(), b"zh") is None
    # unknown language falls back to zh (Whisper.cpp:244-248)
    w2 = pkg.Whisper("micro", util.model_root("micro"), "xx")
    a = util.synth_audio("S", 30000, 9)
    assert w2.run_tokens([a], max_new_tokens=6, honor_eot=False) == whisper.run_tokens([a], max_new_tokens=6, honor_eot=False)
    w2.close()


def test_whisper_cli(tmp_path):
    a = util.synth_audio("S", 67263, 1)
    wav = str(tmp_path / "demo_shape.wav")
    _write_wav(wav, a)
    cli = os.path.join(util.ROOT, "whisper.axera_b200", "whisper_cli")
    out = subprocess.run([cli, "-w", wav, "-t", "micro", "-p", util.model_root("micro"), "--language", "zh"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "Result: " in out.stdout and "RTF: " in out.stdout  # whisper_cli.cpp:102-103


def test_whisper_cli_long_form(whisper, tmp_path):
    # --long (extension): every 30 s window of the file, same text as AX_WHISPER_RunPCMLong on the same samples
    a = np.concatenate([util.synth_audio("S", 480000, 21), util.synth_audio("N", 480000, 22), util.synth_audio("U", 90000, 23)])
    wav = str(tmp_path / "long.wav")
    _write_wav(wav, a)
    pcm16 = (np.clip(a, -1, 1) * 32767).astype("<i2").astype(np.float32) / 32768.0  # what the WAV reader hands to the engine
    expect = whisper.run_long(pcm16)
    cli = os.path.join(util.ROOT, "whisper.axera_b200", "whisper_cli")
    out = subprocess.run([cli, "-w", wav, "-t", "micro", "-p", util.model_root("micro"), "--long"], capture_output=True, text=True, timeout=180)
    assert out.returncode == 0, out.stdout + out.stderr
    line = [l for l in out.stdout.splitlines() if l.startswith("Result: ")][0]
    assert line[len("Result: "):] == expect
    assert len(expect) > 0


def test_concurrent_run_pcm_is_coalesced(whisper):
    """The reference's server calls RunPCM from a thread pool on one non-re-entrant handle (WhisperHTTPServer.hpp:78);
    here concurrent calls are safe and are transcribed together: same text as one at a time, fewer GPU passes than requests."""
    import threading

    audios = [util.synth_audio("NUS"[i % 3], 48000 + 1600 * i, 40 + i) for i in range(12)]
    expect = [whisper.run(a) for a in audios]
    r0, p0 = whisper.stats()
    got = [None] * len(audios)
    errs = []

    def work(i):
        try:
            got[i] = whisper.run(audios[i])
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(audios))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs
    assert got == expect
    r1, p1 = whisper.stats()
    assert r1 - r0 == len(audios)
    assert p1 - p0 < len(audios), "no coalescing happened: %d passes for %d requests" % (p1 - p0, len(audios))
