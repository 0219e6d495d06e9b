#!/usr/bin/env python3
"""Build a whisper.axera-style model directory with seeded random-init weights.

Directory convention is the reference's (/root/reference/cpp/src/Whisper.cpp:87-90,
/root/reference/cpp/src/api/ax_whisper_api.h:33-38) with the two NPU blobs replaced by flat weight files:

    {root}/{type}/{type}-encoder.b200w     encoder.* + decoder.blocks.N.cross_attn.{key,value}.*
    {root}/{type}/{type}-decoder.b200w     the rest of decoder.*
    {root}/{type}/{type}-tokens.txt        "<base64(token bytes)> <rank>\n", line index = id
                                           (/root/reference/model_convert/export_onnx.py:391-417)
    {root}/{type}/{type}_config.json       keys of encoder_meta_data, export_onnx.py:592-625

The split mirrors the reference's encoder graph, which also owns the cross-attention K/V projections
(export_onnx.py:187-213).  Tensor names are openai-whisper state_dict names.  There is no network and
no checkpoint here, so weights are random: numpy PCG64 streams (stable across numpy/torch versions),
recipe in SURVEY.md App. A.6.  Both the CPU oracle and the CUDA engine load these same bytes.

.b200w layout (little endian):
    char[8]  "B200W001"
    u32      n_tensors
    per tensor: u32 name_len, name bytes, u32 dtype (0 = f32), u32 ndim, u64 dims[ndim], u64 byte offset, u64 n_bytes
    data blob, every tensor 256-byte aligned relative to file start
"""
import argparse
import base64
import json
import os
import struct

import numpy as np

ARCHS = {
    #            n_mels d     heads L_enc L_dec n_vocab
    "tiny":  dict(n_mels=80,  d=384,  heads=6,  l_enc=4,  l_dec=4,  n_vocab=51865),
    "base":  dict(n_mels=80,  d=512,  heads=8,  l_enc=6,  l_dec=6,  n_vocab=51865),
    "small": dict(n_mels=80,  d=768,  heads=12, l_enc=12, l_dec=12, n_vocab=51865),
    "turbo": dict(n_mels=128, d=1280, heads=20, l_enc=32, l_dec=4,  n_vocab=51866),
}
# A deliberately tiny architecture for fast CPU tests (not a Whisper release).
ARCHS["micro"] = dict(n_mels=80, d=128, heads=2, l_enc=2, l_dec=2, n_vocab=51865)
ARCH_SEED = {"micro": 7, "tiny": 1, "base": 2, "small": 3, "turbo": 4}
N_AUDIO_CTX = 1500
N_TEXT_CTX = 448

# OpenAI Whisper language codes in tokenizer order (language token id = sot + 1 + index).
LANG_CODES = (
    "en zh de es ru ko fr ja pt tr pl ca nl ar sv it id hi fi vi he uk el ms cs ro da hu ta no th ur hr bg lt la mi "
    "ml cy sk te fa lv bn sr az sl kn et mk br eu is hy ne mn bs kk sq sw gl mr pa si km sn yo so af oc ka be tg sd "
    "gu am yi lo uz fo ht ps tk nn mt sa lb my bo tl mg as tt haw ln ha ba jw su yue"
).split()


def special_tokens(n_vocab):
    """51865-vocab models have 99 languages, 51866-vocab (large-v3/turbo) 100."""
    n_lang = 100 if n_vocab == 51866 else 99
    eot, sot = 50257, 50258
    base = sot + 1 + n_lang
    t = dict(eot=eot, sot=sot, n_lang=n_lang, translate=base, transcribe=base + 1, sot_lm=base + 2,
             sot_prev=base + 3, no_speech=base + 4, no_timestamps=base + 5, timestamp_begin=base + 6, blank_id=220)
    return t


def dims_from_weights(W):
    """Architecture constants of a state dict in openai-whisper naming (head_dim is 64 in every Whisper release)."""
    d = int(W["encoder.conv1.weight"].shape[0])
    n_layers = lambda prefix: 1 + max(int(k.split(".")[2]) for k in W if k.startswith(prefix + ".blocks."))
    return dict(n_mels=int(W["encoder.conv1.weight"].shape[1]), d=d, heads=d // 64, l_enc=n_layers("encoder"),
                l_dec=n_layers("decoder"), n_vocab=int(W["decoder.token_embedding.weight"].shape[0]))


def make_config(arch, dims=None):
    a = dims if dims is not None else ARCHS[arch]
    st = special_tokens(a["n_vocab"])
    codes = LANG_CODES[: st["n_lang"]]
    lang_tokens = [st["sot"] + 1 + i for i in range(st["n_lang"])]
    return {
        "model_type": "whisper-%s" % arch, "version": "1", "maintainer": "k2-fsa",
        "n_mels": a["n_mels"], "n_audio_ctx": N_AUDIO_CTX, "n_audio_state": a["d"], "n_audio_head": a["heads"],
        "n_audio_layer": a["l_enc"], "n_vocab": a["n_vocab"], "n_text_ctx": N_TEXT_CTX, "n_text_state": a["d"],
        "n_text_head": a["heads"], "n_text_layer": a["l_dec"],
        "sot_sequence": "%d,%d,%d" % (st["sot"], st["sot"] + 1, st["transcribe"]),
        "all_language_tokens": ",".join(map(str, lang_tokens)),
        "all_language_codes": ",".join(codes),
        "sot": st["sot"], "sot_index": 0, "eot": st["eot"], "blank_id": st["blank_id"], "is_multilingual": 1,
        "no_speech": st["no_speech"], "non_speech_tokens": "1,2,7,8,9,10,14,25", "transcribe": st["transcribe"],
        "translate": st["translate"], "sot_prev": st["sot_prev"], "sot_lm": st["sot_lm"],
        "no_timestamps": st["no_timestamps"],
    }


def init_weights(arch, seed=None):
    """Random-init recipe. Returns {name: float32 ndarray} in state_dict naming."""
    a = ARCHS[arch]
    d, L_enc, L_dec, V, n_mels = a["d"], a["l_enc"], a["l_dec"], a["n_vocab"], a["n_mels"]
    rng = np.random.default_rng(ARCH_SEED[arch] if seed is None else seed)
    W = {}

    def normal(shape, std):
        return (rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)

    def ln(prefix):
        W[prefix + ".weight"] = (1.0 + normal((d,), 0.02)).astype(np.float32)
        W[prefix + ".bias"] = normal((d,), 0.02)

    def attn(prefix, n_layer, qk_gain=1.2, out_gain=1.0):
        # query/value/out carry a bias, key does not (openai-whisper MultiHeadAttention)
        s = 1.0 / np.sqrt(d)
        W[prefix + ".query.weight"] = normal((d, d), qk_gain * s)
        W[prefix + ".query.bias"] = normal((d,), 0.02)
        W[prefix + ".key.weight"] = normal((d, d), qk_gain * s)
        W[prefix + ".value.weight"] = normal((d, d), s)
        W[prefix + ".value.bias"] = normal((d,), 0.02)
        W[prefix + ".out.weight"] = normal((d, d), out_gain * s / np.sqrt(2.0 * n_layer))
        W[prefix + ".out.bias"] = normal((d,), 0.02)

    def mlp(prefix, n_layer, out_gain=1.0):
        W[prefix + ".0.weight"] = normal((4 * d, d), 1.0 / np.sqrt(d))
        W[prefix + ".0.bias"] = normal((4 * d,), 0.02)
        W[prefix + ".2.weight"] = normal((d, 4 * d), out_gain / np.sqrt(4 * d) / np.sqrt(2.0 * n_layer))
        W[prefix + ".2.bias"] = normal((d,), 0.02)

    W["encoder.conv1.weight"] = normal((d, n_mels, 3), 1.0 / np.sqrt(3 * n_mels))
    W["encoder.conv1.bias"] = normal((d,), 0.02)
    W["encoder.conv2.weight"] = normal((d, d, 3), 1.0 / np.sqrt(3 * d))
    W["encoder.conv2.bias"] = normal((d,), 0.02)
    for i in range(L_enc):
        p = "encoder.blocks.%d" % i
        ln(p + ".attn_ln")
        attn(p + ".attn", L_enc)
        ln(p + ".mlp_ln")
        mlp(p + ".mlp", L_enc)
    ln("encoder.ln_post")

    gain = (0.5 + 1.5 * rng.random((V, 1), dtype=np.float32)).astype(np.float32)
    W["decoder.token_embedding.weight"] = (normal((V, d), 1.0) * gain * np.float32(0.05)).astype(np.float32)
    # Decoder gains were tuned (on the CPU oracle) so that greedy output is not a fixed point: the
    # token-, position- and audio-dependent part of the final hidden state is as large as its constant
    # part (sharp cross-attention, strong MLP and positional terms) -> tens of distinct tokens per 60 steps.
    W["decoder.positional_embedding"] = normal((N_TEXT_CTX, d), 0.1)
    for i in range(L_dec):
        p = "decoder.blocks.%d" % i
        ln(p + ".attn_ln")
        attn(p + ".attn", L_dec)
        ln(p + ".cross_attn_ln")
        attn(p + ".cross_attn", L_dec, qk_gain=2.4, out_gain=0.3)
        ln(p + ".mlp_ln")
        mlp(p + ".mlp", L_dec, out_gain=3.0)
    ln("decoder.ln")
    return W


def is_encoder_file_tensor(name):
    return name.startswith("encoder.") or (".cross_attn.key." in name) or (".cross_attn.value." in name)


def write_b200w(path, tensors):
    names = list(tensors.keys())
    header = bytearray(b"B200W001")
    header += struct.pack("<I", len(names))
    # two passes: header size depends only on names/ndims
    hsize = len(header)
    for n in names:
        t = tensors[n]
        hsize += 4 + len(n.encode()) + 4 + 4 + 8 * t.ndim + 8 + 8
    off = (hsize + 255) // 256 * 256
    offsets = []
    for n in names:
        offsets.append(off)
        off = (off + tensors[n].nbytes + 255) // 256 * 256
    for n, o in zip(names, offsets):
        t = tensors[n]
        nb = n.encode()
        header += struct.pack("<I", len(nb)) + nb + struct.pack("<II", 0, t.ndim)
        header += struct.pack("<%dQ" % t.ndim, *t.shape) + struct.pack("<QQ", o, t.nbytes)
    assert len(header) == hsize
    with open(path, "wb") as f:
        f.write(header)
        for n, o in zip(names, offsets):
            f.seek(o)
            f.write(np.ascontiguousarray(tensors[n], dtype=np.float32).tobytes())
        f.truncate(off)


def read_b200w(path):
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:8] == b"B200W001", "bad magic"
    (n,) = struct.unpack_from("<I", buf, 8)
    pos = 12
    out = {}
    for _ in range(n):
        (nl,) = struct.unpack_from("<I", buf, pos); pos += 4
        name = buf[pos:pos + nl].decode(); pos += nl
        dtype, ndim = struct.unpack_from("<II", buf, pos); pos += 8
        dims = struct.unpack_from("<%dQ" % ndim, buf, pos); pos += 8 * ndim
        off, nbytes = struct.unpack_from("<QQ", buf, pos); pos += 16
        assert dtype == 0
        out[name] = np.frombuffer(buf, dtype=np.float32, count=nbytes // 4, offset=off).reshape(dims)
    return out


def write_tokens(path, tiktoken_path=None, n=50257):
    """Reference format (export_onnx.py:415-417). Without the OpenAI BPE asset (no network) the table is
    synthetic: token i decodes to the UTF-8 string " t<i>" -- decoding random-weight output is noise anyway."""
    with open(path, "w") as f:
        if tiktoken_path:
            for line in open(tiktoken_path):
                if line.strip():
                    tok, rank = line.split()
                    f.write("%s %d\n" % (tok, int(rank)))
        else:
            for i in range(n):
                f.write("%s %d\n" % (base64.b64encode((" t%d" % i).encode()).decode(), i))


def build_model_dir(root, arch, seed=None, tiktoken_path=None, weights=None):
    """weights: optional state dict (openai-whisper names, fp32 numpy); `arch` is then only the directory / file prefix and
    the configuration is derived from the tensor shapes (tools/convert_checkpoint.py)."""
    d = os.path.join(root, arch)
    os.makedirs(d, exist_ok=True)
    W = weights if weights is not None else init_weights(arch, seed)
    # a checkpoint's configuration always comes from its own tensor shapes: `--name turbo` on a large-v3 checkpoint must not
    # silently write the 4-layer turbo table next to 32 decoder layers of weights
    dims = dims_from_weights(W) if weights is not None else None
    enc = {k: v for k, v in W.items() if is_encoder_file_tensor(k)}
    dec = {k: v for k, v in W.items() if not is_encoder_file_tensor(k)}
    write_b200w(os.path.join(d, "%s-encoder.b200w" % arch), enc)
    write_b200w(os.path.join(d, "%s-decoder.b200w" % arch), dec)
    with open(os.path.join(d, "%s_config.json" % arch), "w") as f:
        json.dump(make_config(arch, dims), f, indent=4)
    write_tokens(os.path.join(d, "%s-tokens.txt" % arch), tiktoken_path)
    return d


def load_model_dir(root, arch):
    d = os.path.join(root, arch)
    W = {}
    W.update(read_b200w(os.path.join(d, "%s-encoder.b200w" % arch)))
    W.update(read_b200w(os.path.join(d, "%s-decoder.b200w" % arch)))
    cfg = json.load(open(os.path.join(d, "%s_config.json" % arch)))
    return W, cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", required=True, choices=sorted(ARCHS))
    ap.add_argument("--out", required=True, help="model root; files go to {out}/{arch}/")
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--tiktoken", default=None, help="path to multilingual.tiktoken to convert instead of a synthetic table")
    args = ap.parse_args()
    print(build_model_dir(args.out, args.arch, args.seed, args.tiktoken))


if __name__ == "__main__":
    main()
