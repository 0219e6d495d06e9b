"""The C-ABI shared library loads on a machine without a GPU, exports every declared symbol, and refuses to run without
a B200 (no CPU fallback).  Host-side pieces of the boundary (config parser, WAV reader, base64) are checked through
the library's test hooks."""
import base64
import ctypes
import os
import re
import struct

import numpy as np
import pytest

import make_model
import util


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    declared = set()
    for h in ("ax_whisper_api.h", "b200w_model_abi.h"):
        text = open(os.path.join(util.ROOT, "include", h)).read()
        declared |= set(re.findall(r"\b(AX_WHISPER_[A-Za-z]+|b200w_[a-z0-9_]+)\s*\(", text))
    declared -= {"b200w_engine", "b200w_dims", "b200w_times"}
    assert declared == set(pkg.EXPORTED_SYMBOLS), declared ^ set(pkg.EXPORTED_SYMBOLS)
    for sym in declared:
        assert getattr(lib, sym) is not None


def test_null_argument_conventions(pkg):
    lib = pkg.load_library()
    res = ctypes.c_void_p(123)
    assert lib.AX_WHISPER_RunPCM(None, None, 0, ctypes.byref(res)) == -1   # ax_whisper_api.cpp:143-145
    assert lib.AX_WHISPER_RunFile(None, b"x.wav", ctypes.byref(res)) == -1  # :91-93
    lib.AX_WHISPER_Uninit(None)                                           # no-op, :69-74
    assert lib.AX_WHISPER_Init(None, b".", b"zh") is None


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the behaviour on a machine without a GPU")
def test_no_cpu_fallback(pkg, tmp_path):
    root = util.model_root("micro")
    with pytest.raises(pkg.B200Error, match="no CUDA device|no CPU path"):
        pkg.Engine(root, "micro")
    with pytest.raises(pkg.B200Error):
        pkg.Whisper("micro", root, "zh")


def test_config_parser_reads_reference_format(pkg, tmp_path):
    lib = pkg.load_library()
    root = util.model_root("micro")
    d = pkg.Dims()
    nl = ctypes.c_int()
    assert lib.b200w_test_parse_config(root.encode(), b"micro", ctypes.byref(d), ctypes.byref(nl)) == 0
    a = make_model.ARCHS["micro"]
    assert (d.n_mels, d.n_vocab, d.d_model, d.n_head, d.n_audio_layer, d.n_text_layer) == (80, a["n_vocab"], a["d"], a["heads"], a["l_enc"], a["l_dec"])
    assert (d.sot, d.eot, d.transcribe, d.no_timestamps) == (50258, 50257, 50359, 50363) and nl.value == 99
    t = make_model.make_config("turbo")
    assert (t["n_vocab"], t["transcribe"], t["no_timestamps"], len(t["all_language_codes"].split(","))) == (51866, 50360, 50364, 100)
    # missing file / malformed config -> -1 with a message, never a crash (the reference throws through the C ABI)
    assert lib.b200w_test_parse_config(str(tmp_path).encode(), b"nope", ctypes.byref(d), None) == -1
    os.makedirs(tmp_path / "bad")
    (tmp_path / "bad" / "bad_config.json").write_text('{"n_mels": 80, ')
    assert lib.b200w_test_parse_config(str(tmp_path).encode(), b"bad", ctypes.byref(d), None) == -1
    assert b"config" in lib.b200w_last_error()


def _write_wav(path, data, sr=16000, fmt=1, bits=16):
    n_ch = data.shape[1]
    if fmt == 3:
        raw = data.astype("<f4").tobytes()
    elif bits == 16:
        raw = (data * 32768).clip(-32768, 32767).astype("<i2").tobytes()
    elif bits == 8:
        raw = (data * 128 + 128).clip(0, 255).astype("u1").tobytes()
    elif bits == 32:
        raw = (data.astype(np.float64) * 2147483648).clip(-2147483648, 2147483647).astype("<i4").tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, fmt, n_ch, sr, sr * n_ch * bits // 8, n_ch * bits // 8, bits)
    open(path, "wb").write(hdr + b"LIST" + struct.pack("<I", 4) + b"abcd" + b"data" + struct.pack("<I", len(raw)) + raw)


@pytest.mark.parametrize("fmt,bits,tol", [(1, 16, 1 / 32768), (3, 32, 0), (1, 8, 1 / 128), (1, 32, 1e-6)])
def test_wav_reader(pkg, tmp_path, fmt, bits, tol):
    lib = pkg.load_library()
    rng = np.random.default_rng(0)
    data = rng.uniform(-0.9, 0.9, (1000, 2)).astype(np.float32)
    p = str(tmp_path / "t.wav")
    _write_wav(p, data, fmt=fmt, bits=bits)
    out = np.zeros((1000, 2), np.float32)
    nf, nc, sr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.b200w_test_load_wav(p.encode(), out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 1000, ctypes.byref(nf), ctypes.byref(nc), ctypes.byref(sr)) == 0
    assert (nf.value, nc.value, sr.value) == (1000, 2, 16000)
    assert np.abs(out - data).max() <= tol + 1e-7
    assert lib.b200w_test_load_wav(b"/nonexistent.wav", None, 0, ctypes.byref(nf), ctypes.byref(nc), ctypes.byref(sr)) == -1


def test_reference_demo_wav_shape(pkg):
    demo = "/root/reference/demo.wav"
    if not os.path.exists(demo):
        pytest.skip("reference tree not present")
    lib = pkg.load_library()
    nf, nc, sr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.b200w_test_load_wav(demo.encode(), None, 0, ctypes.byref(nf), ctypes.byref(nc), ctypes.byref(sr)) == 0
    assert (nf.value, nc.value, sr.value) == (67263, 1, 16000)


def test_base64_decoder_is_length_safe(pkg):
    lib = pkg.load_library()
    buf = ctypes.create_string_buffer(256)
    for raw in (b"", b"a", b"ab", b"abc", b" hello", bytes(range(40)), b"\x00nul inside", "擅职".encode()):
        n = lib.b200w_test_base64(base64.b64encode(raw), buf, 256)
        assert n == len(raw) and buf.raw[:n] == raw  # 33-byte and NUL-containing tokens overflow the reference's char[32]
    assert lib.b200w_test_base64(b"!!!!", buf, 256) == 0


def test_tokens_file_format(tmp_path):
    p = str(tmp_path / "tok.txt")
    make_model.write_tokens(p, n=300)
    lines = open(p).read().splitlines()
    assert len(lines) == 300
    tok, rank = lines[42].split(" ")
    assert rank == "42" and base64.b64decode(tok) == b" t42"  # "<base64> <rank>", line index = id (export_onnx.py:415-417)


def _write_aiff(path, data, sr=16000, bits=16, aifc=False):
    """FORM/AIFF (big-endian PCM) or FORM/AIFC with 32-bit floats; the sample rate is an 80-bit extended float."""
    n_frames, n_ch = data.shape
    if aifc:
        raw = data.astype(">f4").tobytes()
        bits = 32
    elif bits == 16:
        raw = (data * 32768).clip(-32768, 32767).astype(">i2").tobytes()
    elif bits == 8:
        raw = (data * 128).clip(-128, 127).astype("i1").tobytes()
    elif bits == 24:
        v = (data.astype(np.float64) * 8388608).clip(-8388608, 8388607).astype(np.int32).reshape(-1)
        raw = b"".join(int(x).to_bytes(3, "big", signed=True) for x in v)
    else:
        raw = (data.astype(np.float64) * 2147483647).clip(-2147483647, 2147483647).astype(">i4").tobytes()
    exp = 16383 + sr.bit_length() - 1
    ext = struct.pack(">HQ", exp, sr << (64 - sr.bit_length()))
    comm = struct.pack(">hIh", n_ch, n_frames, bits) + ext + ((b"fl32" + b"\x00\x00") if aifc else b"")
    ssnd = struct.pack(">II", 0, 0) + raw
    chunks = (b"FVER" + struct.pack(">II", 4, 0xA2805140) if aifc else b"") + b"COMM" + struct.pack(">I", len(comm)) + comm + \
        b"ANNO" + struct.pack(">I", 3) + b"abc\x00" + b"SSND" + struct.pack(">I", len(ssnd)) + ssnd
    open(path, "wb").write(b"FORM" + struct.pack(">I", 4 + len(chunks)) + (b"AIFC" if aifc else b"AIFF") + chunks)


@pytest.mark.parametrize("bits,aifc,tol", [(16, False, 1 / 32768), (8, False, 1 / 128), (24, False, 1 / 8388608), (32, False, 1e-6), (32, True, 0)])
def test_aiff_reader(pkg, tmp_path, bits, aifc, tol):
    """RunFile accepts what AudioFile<float>::load accepts: AIFF / AIFC next to WAV (AudioFile.h:490, :643-770)."""
    lib = pkg.load_library()
    data = np.random.default_rng(1).uniform(-0.9, 0.9, (777, 2)).astype(np.float32)
    p = str(tmp_path / "t.aiff")
    _write_aiff(p, data, bits=bits, aifc=aifc)
    out = np.zeros((777, 2), np.float32)
    nf, nc, sr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.b200w_test_load_wav(p.encode(), out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 777, ctypes.byref(nf), ctypes.byref(nc), ctypes.byref(sr)) == 0, lib.b200w_last_error()
    assert (nf.value, nc.value, sr.value) == (777, 2, 16000)
    assert np.abs(out - data).max() <= tol + 1e-7
    open(p, "wb").write(b"FORM\x00\x00\x00\x04AIFF")  # no COMM / SSND
    assert lib.b200w_test_load_wav(p.encode(), None, 0, ctypes.byref(nf), ctypes.byref(nc), ctypes.byref(sr)) == -1


def _detok(lib, path, ids):
    ids = np.ascontiguousarray(ids, np.int32)
    buf = ctypes.create_string_buffer(1 << 16)
    n = lib.b200w_test_detokenize(path.encode(), ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), len(ids), buf, len(buf))
    assert n >= 0, lib.b200w_last_error()
    return buf.raw[:n]


def test_detokenise_real_tiktoken_vocabulary(pkg, tmp_path):
    """{type}-tokens.txt generated from the reference's own BPE asset (export_onnx.py:391-417 does the same), then every id
    through the library's token-table loader + detokeniser (Whisper.cpp:115-127, :224-229).  The 33-byte token 38538 and the
    NUL token 188 overflow / truncate in the reference's char[32] + strlen path; here they round-trip."""
    asset = "/root/reference/python/assets/multilingual.tiktoken"
    if not os.path.exists(asset):
        pytest.skip("reference tree not present (the GPU box): the committed edge-case fixture is used by test_detokenise_edge_case_fixture")
    lib = pkg.load_library()
    p = str(tmp_path / "real-tokens.txt")
    make_model.write_tokens(p, tiktoken_path=asset)
    table = [base64.b64decode(l.split()[0]) for l in open(asset) if l.strip()]
    assert len(table) == 50257 and len(open(p).read().splitlines()) == 50257
    assert table[188] == b"\x00" and len(table[38538]) == 33
    for lo in range(0, 50257, 4096):  # every id, in runs (the hook concatenates)
        ids = list(range(lo, min(lo + 4096, 50257)))
        assert _detok(lib, p, ids) == b"".join(table[i] for i in ids)
    seq = [38538, 188, 38538, 50257, 50363, 51864, 220, 188]  # specials (>= 50257) carry no text and are skipped
    assert _detok(lib, p, seq) == table[38538] + b"\x00" + table[38538] + table[220] + b"\x00"
    assert _detok(lib, p, [-1, 10 ** 6]) == b""


def test_detokenise_edge_case_fixture(pkg, tmp_path):
    """Same check from the committed fixture (tests/golden/tiktoken_edge_cases.json, written by tools/make_golden.py from the
    reference's multilingual.tiktoken): the longest tokens, the NUL token, multi-byte UTF-8 fragments."""
    import json
    fx = json.load(open(os.path.join(util.ROOT, "tests", "golden", "tiktoken_edge_cases.json")))
    lib = pkg.load_library()
    ids = sorted(int(k) for k in fx["tokens"])
    n = max(ids) + 1
    p = str(tmp_path / "edge-tokens.txt")
    with open(p, "w") as f:
        for i in range(n):
            f.write("%s %d\n" % (fx["tokens"].get(str(i), base64.b64encode((" t%d" % i).encode()).decode()), i))
    want = b"".join(base64.b64decode(fx["tokens"][str(i)]) for i in ids)
    assert _detok(lib, p, ids) == want
    assert any(len(base64.b64decode(v)) == 33 for v in fx["tokens"].values()) and any(b"\x00" in base64.b64decode(v) for v in fx["tokens"].values())
