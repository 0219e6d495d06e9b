"""whisper.axera_b200 -- Python host binding of libax_whisper.so (B200 / sm_100a build).

The product is the C++/CUDA library next to this file; Python is only the host-side mirror used by the tests,
bench.py and scripts.  Two levels, like the reference:

  * `Whisper`  mirrors the reference's Python host class (/root/reference/python/whisper.py:24-57, run :213-271)
               and the C API it wraps (/root/reference/cpp/src/api/ax_whisper_api.h): Whisper(model_type,
               model_path, language).run(audio) -> text.
  * `Engine`   the model-level ABI (include/b200w_model_abi.h): logmel / encoder / decoder_main / decoder_loop /
               greedy / transcribe, numpy in, numpy out.

There is no CPU fallback: loading fails loudly if the shared library is missing, and creating an engine fails
loudly if no B200 is visible.  The directory name contains a dot, so import it by path
(see __graft_entry__.load_package()).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200W_LIB: another build of the same library (A/B measurements of two revisions on one GPU box, scripts/ab_build.sh)
LIB_PATH = os.environ.get("B200W_LIB") or os.path.join(_HERE, "libax_whisper.so")
N_FRAMES = 3000
N_AUDIO_CTX = 1500
N_TEXT_CTX = 448
CHUNK_SAMPLES = 480000

_c_float_p = ctypes.POINTER(ctypes.c_float)
_c_int_p = ctypes.POINTER(ctypes.c_int)


class Dims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "n_mels", "n_vocab", "d_model", "n_head", "n_audio_layer", "n_text_layer", "n_audio_ctx", "n_text_ctx",
        "sot", "eot", "transcribe", "no_timestamps")]


class Times(ctypes.Structure):
    _fields_ = [("h2d_ms", ctypes.c_float), ("mel_ms", ctypes.c_float), ("encoder_ms", ctypes.c_float),
                ("decode_ms", ctypes.c_float), ("total_ms", ctypes.c_float), ("decode_steps", ctypes.c_int),
                ("kernel_launches", ctypes.c_longlong)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/ax_whisper_api.h and include/b200w_model_abi.h declare
EXPORTED_SYMBOLS = [
    "AX_WHISPER_Init", "AX_WHISPER_Uninit", "AX_WHISPER_RunFile", "AX_WHISPER_RunPCM", "AX_WHISPER_RunPCMBatch",
    "AX_WHISPER_RunPCMTokens", "AX_WHISPER_RunPCMLong", "AX_WHISPER_GetStats", "AX_WHISPER_LastError",
    "b200w_last_error", "b200w_engine_create", "b200w_engine_destroy", "b200w_get_dims", "b200w_sot_sequence",
    "b200w_logmel", "b200w_encoder", "b200w_decoder_main", "b200w_decoder_loop", "b200w_greedy", "b200w_transcribe",
    "b200w_upload_pcm", "b200w_transcribe_resident", "b200w_time_stage", "b200w_selftest_gemm", "b200w_mel_tables",
    "b200w_get_cross_kv", "b200w_set_cross_kv", "b200w_set_self_kv", "b200w_get_self_kv", "b200w_decoder_step", "b200w_set_logit_rows", "b200w_decode_stats",
    "b200w_test_parse_config", "b200w_test_load_wav", "b200w_test_base64", "b200w_test_detokenize", "b200w_selftest_attention", "b200w_selftest_cross_attention",
]

_lib = None


def load_library():
    """dlopen libax_whisper.so (built by __graft_entry__.build() / csrc/build.sh). Raises if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, cp, ci, cl = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_long
    lib.AX_WHISPER_Init.restype = vp
    lib.AX_WHISPER_Init.argtypes = [cp, cp, cp]
    lib.AX_WHISPER_Uninit.restype = None
    lib.AX_WHISPER_Uninit.argtypes = [vp]
    lib.AX_WHISPER_RunFile.argtypes = [vp, cp, ctypes.POINTER(vp)]
    lib.AX_WHISPER_RunPCM.argtypes = [vp, _c_float_p, ci, ctypes.POINTER(vp)]
    lib.AX_WHISPER_RunPCMBatch.argtypes = [vp, ctypes.POINTER(_c_float_p), _c_int_p, ci, ctypes.POINTER(vp)]
    lib.AX_WHISPER_RunPCMTokens.argtypes = [vp, ctypes.POINTER(_c_float_p), _c_int_p, ci, ci, ci, _c_int_p, ci, _c_int_p]
    lib.AX_WHISPER_RunPCMLong.argtypes = [vp, _c_float_p, cl, ci, ctypes.POINTER(vp)]
    lib.AX_WHISPER_GetStats.argtypes = [vp, ctypes.POINTER(cl), ctypes.POINTER(cl)]
    lib.AX_WHISPER_LastError.restype = cp
    lib.b200w_last_error.restype = cp
    lib.b200w_engine_create.argtypes = [cp, cp, ci, ci, ctypes.POINTER(vp)]
    lib.b200w_engine_destroy.restype = None
    lib.b200w_engine_destroy.argtypes = [vp]
    lib.b200w_get_dims.argtypes = [vp, ctypes.POINTER(Dims)]
    lib.b200w_sot_sequence.argtypes = [vp, cp, _c_int_p]
    lib.b200w_logmel.argtypes = [vp, _c_float_p, cl, _c_int_p, ci, _c_float_p]
    lib.b200w_encoder.argtypes = [vp, _c_float_p, ci, _c_float_p, _c_float_p]
    lib.b200w_decoder_main.argtypes = [vp, _c_int_p, ci, ci, _c_float_p, _c_float_p, _c_float_p]
    lib.b200w_decoder_loop.argtypes = [vp, _c_int_p, ci, ci, _c_float_p, _c_float_p, _c_float_p]
    _sig(lib, "b200w_get_cross_kv", [vp, ci, ci, _c_float_p, _c_float_p])
    _sig(lib, "b200w_set_cross_kv", [vp, _c_float_p, _c_float_p, ci])
    _sig(lib, "b200w_set_self_kv", [vp, _c_float_p, _c_float_p, ci, ci])
    _sig(lib, "b200w_get_self_kv", [vp, _c_float_p, _c_float_p, ci, ci])
    _sig(lib, "b200w_decoder_step", [vp, _c_int_p, _c_float_p, _c_float_p, _c_float_p, _c_float_p, ci, _c_int_p, ci, _c_float_p, _c_float_p, _c_float_p])
    _sig(lib, "b200w_set_logit_rows", [vp, _c_int_p, ci])
    _sig(lib, "b200w_decode_stats", [vp, ctypes.POINTER(cl), _c_int_p])
    lib.b200w_greedy.argtypes = [vp, ci, cp, ci, ci, _c_int_p, ci, _c_float_p, _c_int_p, ci, _c_int_p]
    lib.b200w_transcribe.argtypes = [vp, _c_float_p, cl, _c_int_p, ci, cp, ci, ci, _c_int_p, ci, _c_int_p, ctypes.POINTER(Times)]
    lib.b200w_upload_pcm.argtypes = [vp, _c_float_p, cl, _c_int_p, ci]
    lib.b200w_transcribe_resident.argtypes = [vp, ci, cp, ci, ci, _c_int_p, ci, _c_int_p, ctypes.POINTER(Times)]
    lib.b200w_time_stage.argtypes = [vp, ci, ci, ci, ci, _c_float_p]
    lib.b200w_selftest_gemm.argtypes = [ci, ci, ci, ci, ci, ctypes.c_uint, _c_float_p, _c_float_p]
    lib.b200w_mel_tables.argtypes = [ci, _c_float_p, _c_float_p]
    lib.b200w_selftest_attention.argtypes = [ci, ci, ci, ctypes.c_uint, _c_float_p, _c_float_p]
    lib.b200w_selftest_cross_attention.argtypes = [ci, ci, ci, ctypes.c_uint, _c_float_p, _c_float_p]
    lib.b200w_test_parse_config.argtypes = [cp, cp, ctypes.POINTER(Dims), _c_int_p]
    lib.b200w_test_load_wav.argtypes = [cp, _c_float_p, ci, _c_int_p, _c_int_p, _c_int_p]
    lib.b200w_test_base64.argtypes = [cp, ctypes.c_char_p, ci]
    _sig(lib, "b200w_test_detokenize", [cp, _c_int_p, ci, ctypes.c_char_p, ci])
    _lib = lib
    return lib


def _sig(lib, name, argtypes):
    """argtypes of an entry point added after round 1; an older A/B build (B200W_LIB) may lack it."""
    try:
        getattr(lib, name).argtypes = argtypes
    except AttributeError:
        if not os.environ.get("B200W_LIB"):
            raise


def _fp(a):
    return a.ctypes.data_as(_c_float_p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(_c_int_p) if a is not None else None


class B200Error(RuntimeError):
    pass


class Engine:
    """Model-level ABI (include/b200w_model_abi.h)."""

    def __init__(self, model_path, model_type, device=0, max_batch=1):
        self.lib = load_library()
        h = ctypes.c_void_p()
        if self.lib.b200w_engine_create(str(model_path).encode(), model_type.encode(), device, max_batch, ctypes.byref(h)) != 0:
            raise B200Error(self.lib.b200w_last_error().decode())
        self.h = h
        d = Dims()
        self._check(self.lib.b200w_get_dims(self.h, ctypes.byref(d)))
        self.dims = d

    def _check(self, rc):
        if rc != 0:
            raise B200Error(self.lib.b200w_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200w_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sot_sequence(self, language="zh"):
        out = np.zeros(4, np.int32)
        self._check(self.lib.b200w_sot_sequence(self.h, language.encode(), _ip(out)))
        return out.tolist()

    @staticmethod
    def _pack_pcm(audios):
        if isinstance(audios, np.ndarray) and audios.ndim == 2:
            pcm = np.ascontiguousarray(audios, np.float32)
            n = np.full(pcm.shape[0], pcm.shape[1], np.int32)
            return pcm, n
        n = np.array([len(a) for a in audios], np.int32)
        pcm = np.zeros((len(audios), int(n.max())), np.float32)
        for i, a in enumerate(audios):
            pcm[i, : len(a)] = a
        return pcm, n

    def logmel(self, audios):
        """list of 1-D float32 arrays (or [B, n] array) -> [B, n_mels, 3000] float32."""
        pcm, n = self._pack_pcm(audios)
        out = np.empty((pcm.shape[0], self.dims.n_mels, N_FRAMES), np.float32)
        self._check(self.lib.b200w_logmel(self.h, _fp(pcm), pcm.shape[1], _ip(n), pcm.shape[0], _fp(out)))
        return out

    def encoder(self, mel=None, batch=None, return_cross=True):
        """mel [B, n_mels, 3000] (or None to use the resident log-mel of the last logmel() call)."""
        if mel is not None:
            mel = np.ascontiguousarray(mel, np.float32)
            batch = mel.shape[0]
        L, d = self.dims.n_text_layer, self.dims.d_model
        ck = cv = None
        if return_cross:
            ck = np.empty((L, batch, N_AUDIO_CTX, d), np.float32)
            cv = np.empty((L, batch, N_AUDIO_CTX, d), np.float32)
        self._check(self.lib.b200w_encoder(self.h, _fp(mel), batch, _fp(ck), _fp(cv)))
        return ck, cv

    def decoder_main(self, sot_tokens, batch):
        L, d, V = self.dims.n_text_layer, self.dims.d_model, self.dims.n_vocab
        toks = np.asarray(sot_tokens, np.int32)
        logits = np.empty((batch, V), np.float32)
        k = np.empty((L, batch, len(toks), d), np.float32)
        v = np.empty((L, batch, len(toks), d), np.float32)
        self._check(self.lib.b200w_decoder_main(self.h, _ip(toks), len(toks), batch, _fp(logits), _fp(k), _fp(v)))
        return logits, k, v

    def decoder_loop(self, tokens, offset):
        L, d, V = self.dims.n_text_layer, self.dims.d_model, self.dims.n_vocab
        toks = np.ascontiguousarray(tokens, np.int32)
        B = len(toks)
        logits = np.empty((B, V), np.float32)
        k = np.empty((L, B, d), np.float32)
        v = np.empty((L, B, d), np.float32)
        self._check(self.lib.b200w_decoder_loop(self.h, _ip(toks), int(offset), B, _fp(logits), _fp(k), _fp(v)))
        return logits, k, v

    def get_cross_kv(self, b0, nb):
        """resident cross K/V of sequences [b0, b0 + nb) as f32 [L, nb, 1500, d] (the reference encoder's out0 / out1)."""
        L, d = self.dims.n_text_layer, self.dims.d_model
        ck = np.empty((L, nb, N_AUDIO_CTX, d), np.float32)
        cv = np.empty((L, nb, N_AUDIO_CTX, d), np.float32)
        self._check(self.lib.b200w_get_cross_kv(self.h, b0, nb, _fp(ck), _fp(cv)))
        return ck, cv

    def set_cross_kv(self, cross_k, cross_v):
        ck = np.ascontiguousarray(cross_k, np.float32)
        cv = np.ascontiguousarray(cross_v, np.float32)
        assert ck.shape == cv.shape and ck.shape[0] == self.dims.n_text_layer and ck.shape[2:] == (N_AUDIO_CTX, self.dims.d_model)
        self._check(self.lib.b200w_set_cross_kv(self.h, _fp(ck), _fp(cv), ck.shape[1]))

    def set_self_kv(self, self_k, self_v, n_valid):
        sk = np.ascontiguousarray(self_k, np.float32)
        sv = np.ascontiguousarray(self_v, np.float32)
        assert sk.shape == sv.shape and sk.shape[2:] == (N_TEXT_CTX, self.dims.d_model)
        self._check(self.lib.b200w_set_self_kv(self.h, _fp(sk), _fp(sv), int(n_valid), sk.shape[1]))

    def get_self_kv(self, batch, n_rows):
        L, d = self.dims.n_text_layer, self.dims.d_model
        sk = np.empty((L, batch, n_rows, d), np.float32)
        sv = np.empty((L, batch, n_rows, d), np.float32)
        self._check(self.lib.b200w_get_self_kv(self.h, _fp(sk), _fp(sv), n_rows, batch))
        return sk, sv

    def decoder_step(self, tokens, offset, self_k=None, self_v=None, cross_k=None, cross_v=None, mask=None):
        """The reference decoder graph's stateless call: every cache may be passed in (None = keep the resident one)."""
        L, d, V = self.dims.n_text_layer, self.dims.d_model, self.dims.n_vocab
        toks = np.ascontiguousarray(tokens, np.int32)
        B = len(toks)
        f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        sk, sv, ck, cv = f(self_k), f(self_v), f(cross_k), f(cross_v)
        m = None if mask is None else np.ascontiguousarray(mask, np.int32)
        logits = np.empty((B, V), np.float32)
        k = np.empty((L, B, d), np.float32)
        v = np.empty((L, B, d), np.float32)
        self._check(self.lib.b200w_decoder_step(self.h, _ip(toks), _fp(sk), _fp(sv), _fp(ck), _fp(cv), int(offset), _ip(m), B, _fp(logits),
                                                _fp(k), _fp(v)))
        return logits, k, v

    def greedy(self, batch, language="zh", max_new_tokens=0, honor_eot=True, forced_tokens=None, keep_logits=False, logit_rows=None):
        """logit_rows: with keep_logits, only these sequences' logits are returned ([n_steps, len(logit_rows), V])."""
        forced = None
        flen = 0
        if forced_tokens is not None:
            forced = np.ascontiguousarray(forced_tokens, np.int32)
            flen = forced.shape[1]
        n_max = max_new_tokens if max_new_tokens > 0 else N_TEXT_CTX - 4
        rows = np.ascontiguousarray(logit_rows if logit_rows is not None else [], np.int32)
        self._check(self.lib.b200w_set_logit_rows(self.h, _ip(rows), len(rows)))
        logits = np.zeros((n_max, len(rows) if len(rows) else batch, self.dims.n_vocab), np.float32) if keep_logits else None
        toks = np.zeros((batch, N_TEXT_CTX), np.int32)
        n = np.zeros(batch, np.int32)
        self._check(self.lib.b200w_greedy(self.h, batch, language.encode(), max_new_tokens, int(honor_eot), _ip(forced), flen,
                                          _fp(logits), _ip(toks), N_TEXT_CTX, _ip(n)))
        return [toks[b, : n[b]].tolist() for b in range(batch)], logits

    def decode_stats(self):
        """(EOT compactions of the slot list since creation, sequences still decoding when the last greedy loop stopped)."""
        c, a = ctypes.c_long(), ctypes.c_int()
        self._check(self.lib.b200w_decode_stats(self.h, ctypes.byref(c), ctypes.byref(a)))
        return c.value, a.value

    def transcribe(self, audios, language="zh", max_new_tokens=0, honor_eot=True):
        pcm, n = self._pack_pcm(audios)
        B = pcm.shape[0]
        toks = np.zeros((B, N_TEXT_CTX), np.int32)
        nt = np.zeros(B, np.int32)
        t = Times()
        self._check(self.lib.b200w_transcribe(self.h, _fp(pcm), pcm.shape[1], _ip(n), B, language.encode(), max_new_tokens,
                                              int(honor_eot), _ip(toks), N_TEXT_CTX, _ip(nt), ctypes.byref(t)))
        return [toks[b, : nt[b]].tolist() for b in range(B)], t.as_dict()

    def upload_pcm(self, audios):
        pcm, n = self._pack_pcm(audios)
        self._check(self.lib.b200w_upload_pcm(self.h, _fp(pcm), pcm.shape[1], _ip(n), pcm.shape[0]))
        return pcm.shape[0]

    def transcribe_resident(self, batch, language="zh", max_new_tokens=0, honor_eot=True):
        toks = np.zeros((batch, N_TEXT_CTX), np.int32)
        nt = np.zeros(batch, np.int32)
        t = Times()
        self._check(self.lib.b200w_transcribe_resident(self.h, batch, language.encode(), max_new_tokens, int(honor_eot), _ip(toks),
                                                       N_TEXT_CTX, _ip(nt), ctypes.byref(t)))
        return [toks[b, : nt[b]].tolist() for b in range(batch)], t.as_dict()

    def time_stage(self, stage, batch, iters=1, n_steps=228):
        """CUDA-event time (ms) of `iters` runs of one stage on resident data: 0 log-mel, 1 encoder, 2 decode."""
        ms = ctypes.c_float()
        self._check(self.lib.b200w_time_stage(self.h, stage, batch, iters, n_steps, ctypes.byref(ms)))
        return ms.value


class Whisper:
    """Reference-shaped front door: Whisper(model_type, model_path, language).run(audio) -> str
    (/root/reference/python/whisper.py:24-57,213-271 over the C API of /root/reference/cpp/src/api/ax_whisper_api.h)."""

    def __init__(self, model_type, model_path, language="zh"):
        self.lib = load_library()
        self.h = self.lib.AX_WHISPER_Init(model_type.encode(), str(model_path).encode(), language.encode())
        if not self.h:
            raise B200Error("AX_WHISPER_Init failed: " + self.lib.AX_WHISPER_LastError().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.AX_WHISPER_Uninit(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _take(self, p):
        s = ctypes.string_at(p).decode("utf-8", errors="replace")
        ctypes.CDLL(None).free(ctypes.c_void_p(p))
        return s

    def run(self, audio):
        """audio: path to a WAV file, or 1-D float32 PCM at 16 kHz."""
        res = ctypes.c_void_p()
        if isinstance(audio, (str, os.PathLike)):
            rc = self.lib.AX_WHISPER_RunFile(self.h, os.fspath(audio).encode(), ctypes.byref(res))
        else:
            a = np.ascontiguousarray(audio, np.float32)
            rc = self.lib.AX_WHISPER_RunPCM(self.h, _fp(a), len(a), ctypes.byref(res))
        if rc != 0:
            raise B200Error("AX_WHISPER_Run failed: " + self.lib.AX_WHISPER_LastError().decode())
        return self._take(res.value)

    def stats(self):
        """(single-utterance requests served, GPU passes they took): concurrent run() calls are coalesced into batches."""
        a, b = ctypes.c_long(), ctypes.c_long()
        self.lib.AX_WHISPER_GetStats(self.h, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    def run_long(self, audio, window_batch=0):
        """Long-form: consecutive 30 s windows transcribed as one batch, texts concatenated (AX_WHISPER_RunPCMLong)."""
        a = np.ascontiguousarray(audio, np.float32)
        res = ctypes.c_void_p()
        if self.lib.AX_WHISPER_RunPCMLong(self.h, _fp(a), len(a), window_batch, ctypes.byref(res)) != 0:
            raise B200Error("AX_WHISPER_RunPCMLong failed: " + self.lib.AX_WHISPER_LastError().decode())
        return self._take(res.value)

    def run_tokens(self, audios, max_new_tokens=0, honor_eot=True):
        arrs = [np.ascontiguousarray(a, np.float32) for a in audios]
        B = len(arrs)
        ptrs = (_c_float_p * B)(*[_fp(a) for a in arrs])
        n = np.array([len(a) for a in arrs], np.int32)
        toks = np.zeros((B, N_TEXT_CTX), np.int32)
        nt = np.zeros(B, np.int32)
        rc = self.lib.AX_WHISPER_RunPCMTokens(self.h, ptrs, _ip(n), B, max_new_tokens, int(honor_eot), _ip(toks), N_TEXT_CTX, _ip(nt))
        if rc != 0:
            raise B200Error("AX_WHISPER_RunPCMTokens failed: " + self.lib.AX_WHISPER_LastError().decode())
        return [toks[b, : nt[b]].tolist() for b in range(B)]


def selftest_gemm(M, N, K, block_n, epilogue, seed=0):
    lib = load_library()
    d, r = ctypes.c_float(), ctypes.c_float()
    if lib.b200w_selftest_gemm(M, N, K, block_n, epilogue, seed, ctypes.byref(d), ctypes.byref(r)) != 0:
        raise B200Error(lib.b200w_last_error().decode())
    return d.value, r.value


def selftest_attention(B, T, n_head, seed=0):
    lib = load_library()
    d, r = ctypes.c_float(), ctypes.c_float()
    if lib.b200w_selftest_attention(B, T, n_head, seed, ctypes.byref(d), ctypes.byref(r)) != 0:
        raise B200Error(lib.b200w_last_error().decode())
    return d.value, r.value


def selftest_cross_attention(B, n_head, T, seed=0):
    lib = load_library()
    d, r = ctypes.c_float(), ctypes.c_float()
    if lib.b200w_selftest_cross_attention(B, n_head, T, seed, ctypes.byref(d), ctypes.byref(r)) != 0:
        raise B200Error(lib.b200w_last_error().decode())
    return d.value, r.value


def mel_tables(n_mels):
    lib = load_library()
    bank = np.zeros((n_mels, 201), np.float32)
    win = np.zeros(400, np.float32)
    if lib.b200w_mel_tables(n_mels, _fp(bank), _fp(win)) != 0:
        raise B200Error("mel_tables failed")
    return bank, win
