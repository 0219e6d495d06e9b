#!/usr/bin/env python3
"""Turn a trained Whisper checkpoint into a model directory the B200 engine (and the oracle) load:
{out}/{name}/{name}-encoder.b200w, -decoder.b200w, -tokens.txt, _config.json -- the layout of the reference's
{type}-encoder.axmodel / -decoder.axmodel / -tokens.txt / _config.json (/root/reference/cpp/src/Whisper.cpp:87-90,
config keys of /root/reference/model_convert/export_onnx.py:592-625).

  --openai small.pt          an openai-whisper checkpoint ({"dims", "model_state_dict"}; tensor names are used as they are)
  --hf DIR                   a local transformers WhisperForConditionalGeneration directory (names mapped below)
  --tiktoken FILE            OpenAI's multilingual.tiktoken (the reference ships it as python/assets/multilingual.tiktoken);
                             without it a synthetic token table is written and only token ids are meaningful

Weights are stored as fp32; the engine rounds the matrices to bf16 when it loads them."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_model


def from_openai_state_dict(sd):
    """openai-whisper names are the engine's names; drop buffers the engine rebuilds (sinusoidal positions, causal mask)."""
    skip = ("encoder.positional_embedding", "decoder.mask")
    return {k: np.asarray(v.detach().float().cpu().numpy() if hasattr(v, "detach") else v, np.float32)
            for k, v in sd.items() if k not in skip}


def from_hf_state_dict(sd):
    """transformers' WhisperForConditionalGeneration -> openai-whisper names (the inverse of the mapping that
    tests/test_oracle_vs_transformers.py uses to validate the oracle)."""
    g = lambda k: np.asarray(sd[k].detach().float().cpu().numpy(), np.float32)
    W = {}
    for c in ("conv1", "conv2"):
        W["encoder.%s.weight" % c], W["encoder.%s.bias" % c] = g("model.encoder.%s.weight" % c), g("model.encoder.%s.bias" % c)

    def attn(src, dst):
        for hf, ours in (("q_proj", "query"), ("k_proj", "key"), ("v_proj", "value"), ("out_proj", "out")):
            W[dst + "." + ours + ".weight"] = g(src + "." + hf + ".weight")
            if ours != "key":  # Whisper's key projection has no bias
                W[dst + "." + ours + ".bias"] = g(src + "." + hf + ".bias")

    def ln(src, dst):
        W[dst + ".weight"], W[dst + ".bias"] = g(src + ".weight"), g(src + ".bias")

    def mlp(src, dst):
        for hf, ours in (("fc1", "mlp.0"), ("fc2", "mlp.2")):
            W[dst + "." + ours + ".weight"], W[dst + "." + ours + ".bias"] = g(src + "." + hf + ".weight"), g(src + "." + hf + ".bias")

    n_enc = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("model.encoder.layers."))
    n_dec = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("model.decoder.layers."))
    for i in range(n_enc):
        s, d = "model.encoder.layers.%d" % i, "encoder.blocks.%d" % i
        attn(s + ".self_attn", d + ".attn")
        ln(s + ".self_attn_layer_norm", d + ".attn_ln")
        ln(s + ".final_layer_norm", d + ".mlp_ln")
        mlp(s, d)
    ln("model.encoder.layer_norm", "encoder.ln_post")
    W["decoder.token_embedding.weight"] = g("model.decoder.embed_tokens.weight")
    W["decoder.positional_embedding"] = g("model.decoder.embed_positions.weight")
    for i in range(n_dec):
        s, d = "model.decoder.layers.%d" % i, "decoder.blocks.%d" % i
        attn(s + ".self_attn", d + ".attn")
        attn(s + ".encoder_attn", d + ".cross_attn")
        ln(s + ".self_attn_layer_norm", d + ".attn_ln")
        ln(s + ".encoder_attn_layer_norm", d + ".cross_attn_ln")
        ln(s + ".final_layer_norm", d + ".mlp_ln")
        mlp(s, d)
    ln("model.decoder.layer_norm", "decoder.ln")
    return W


def check(W):
    dims = make_model.dims_from_weights(W)
    if dims["d"] % 128 or dims["d"] > 1280:
        raise SystemExit("d_model %d: the engine's kernels cover multiples of 128 up to 1280" % dims["d"])
    if dims["n_vocab"] not in (51865, 51866):
        raise SystemExit("n_vocab %d: only the multilingual vocabularies (51865, 51866) have a special-token table here" % dims["n_vocab"])
    if W["decoder.positional_embedding"].shape[0] != make_model.N_TEXT_CTX:
        raise SystemExit("n_text_ctx must be 448")
    return dims


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    src = ap.add_mutually_exclusive_group(required=True)
    src.add_argument("--openai")
    src.add_argument("--hf")
    ap.add_argument("--name", required=True, help="model_type: directory and file prefix (what AX_WHISPER_Init receives)")
    ap.add_argument("--out", required=True, help="model_path: files go to {out}/{name}/")
    ap.add_argument("--tiktoken", default=None)
    args = ap.parse_args()
    import torch

    if args.openai:
        ck = torch.load(args.openai, map_location="cpu", weights_only=True)
        W = from_openai_state_dict(ck["model_state_dict"] if "model_state_dict" in ck else ck)
    else:
        import transformers

        W = from_hf_state_dict(transformers.WhisperForConditionalGeneration.from_pretrained(args.hf, local_files_only=True).state_dict())
    dims = check(W)
    print(make_model.build_model_dir(args.out, args.name, tiktoken_path=args.tiktoken, weights=W), dims)


if __name__ == "__main__":
    main()
