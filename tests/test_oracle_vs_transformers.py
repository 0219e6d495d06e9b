"""The torch restatement (oracle/whisper_oracle.py) against an independent implementation of the same architecture:
transformers' WhisperForConditionalGeneration with our seeded weights mapped in.  fp32 on CPU, agreement <= 2e-4.
(The reference's own stack -- openai-whisper + onnxruntime -- is not installed in this image)"""
import numpy as np
import pytest
import torch

import make_model
import util

transformers = pytest.importorskip("transformers")


def _hf_model(arch, W):
    a = make_model.ARCHS[arch]
    cfg = transformers.WhisperConfig(
        vocab_size=a["n_vocab"], num_mel_bins=a["n_mels"], d_model=a["d"], encoder_layers=a["l_enc"], decoder_layers=a["l_dec"],
        encoder_attention_heads=a["heads"], decoder_attention_heads=a["heads"], encoder_ffn_dim=4 * a["d"], decoder_ffn_dim=4 * a["d"],
        max_source_positions=1500, max_target_positions=448, activation_function="gelu", dropout=0.0, attention_dropout=0.0,
        activation_dropout=0.0, scale_embedding=False)
    m = transformers.WhisperForConditionalGeneration(cfg).eval()
    sd = {}
    t = lambda k: torch.from_numpy(np.array(W[k]))

    def attn(src, dst):
        for ours, hf in (("query", "q_proj"), ("key", "k_proj"), ("value", "v_proj"), ("out", "out_proj")):
            sd[dst + "." + hf + ".weight"] = t(src + "." + ours + ".weight")
            if ours != "key":
                sd[dst + "." + hf + ".bias"] = t(src + "." + ours + ".bias")

    def ln(src, dst):
        sd[dst + ".weight"], sd[dst + ".bias"] = t(src + ".weight"), t(src + ".bias")

    for c in ("conv1", "conv2"):
        sd["model.encoder.%s.weight" % c], sd["model.encoder.%s.bias" % c] = t("encoder.%s.weight" % c), t("encoder.%s.bias" % c)
    import whisper_oracle
    sd["model.encoder.embed_positions.weight"] = whisper_oracle.sinusoids(1500, a["d"])
    for i in range(a["l_enc"]):
        s, d = "encoder.blocks.%d" % i, "model.encoder.layers.%d" % i
        attn(s + ".attn", d + ".self_attn")
        ln(s + ".attn_ln", d + ".self_attn_layer_norm")
        ln(s + ".mlp_ln", d + ".final_layer_norm")
        for ours, hf in (("mlp.0", "fc1"), ("mlp.2", "fc2")):
            sd[d + "." + hf + ".weight"], sd[d + "." + hf + ".bias"] = t(s + "." + ours + ".weight"), t(s + "." + ours + ".bias")
    ln("encoder.ln_post", "model.encoder.layer_norm")
    sd["model.decoder.embed_tokens.weight"] = t("decoder.token_embedding.weight")
    sd["model.decoder.embed_positions.weight"] = t("decoder.positional_embedding")
    for i in range(a["l_dec"]):
        s, d = "decoder.blocks.%d" % i, "model.decoder.layers.%d" % i
        attn(s + ".attn", d + ".self_attn")
        attn(s + ".cross_attn", d + ".encoder_attn")
        ln(s + ".attn_ln", d + ".self_attn_layer_norm")
        ln(s + ".cross_attn_ln", d + ".encoder_attn_layer_norm")
        ln(s + ".mlp_ln", d + ".final_layer_norm")
        for ours, hf in (("mlp.0", "fc1"), ("mlp.2", "fc2")):
            sd[d + "." + hf + ".weight"], sd[d + "." + hf + ".bias"] = t(s + "." + ours + ".weight"), t(s + "." + ours + ".bias")
    ln("decoder.ln", "model.decoder.layer_norm")
    sd["proj_out.weight"] = t("decoder.token_embedding.weight")
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("k_proj.bias" in k for k in missing), missing
    return m


# micro: the CPU-test architecture; tiny: a Whisper release (d = 384); small: the headline architecture (d = 768, 12 + 12 layers),
# one chunk.  fp32 on CPU both sides: agreement <= 2e-4 on every architecture (measured ~2e-6 on the encoder output).
@pytest.mark.parametrize("arch,n_chunks,tol", [("micro", 2, 2e-4), ("tiny", 2, 2e-4), ("small", 1, 2e-4)])
def test_oracle_matches_transformers(arch, n_chunks, tol):
    torch.set_num_threads(8)
    W = make_model.init_weights(arch)
    cfg = make_model.make_config(arch)
    import whisper_oracle
    o = whisper_oracle.Oracle(W, cfg)
    hf = _hf_model(arch, W)
    rng = np.random.default_rng(3)
    mel = (rng.random((n_chunks, cfg["n_mels"], 3000), dtype=np.float32) * 2 - 1)
    with torch.no_grad():
        xa = o.audio_features(mel)
        hf_xa = hf.model.encoder(torch.from_numpy(mel)).last_hidden_state
        enc_err = float((xa - hf_xa).abs().max())
        print("%s encoder output: max-abs diff vs transformers %.2e (max |x| %.2f)" % (arch, enc_err, float(xa.abs().max())))
        assert enc_err <= tol
        ck, cv = o.encoder(mel)
        r = o.greedy(ck, cv, max_new_tokens=6, honor_eot=False, keep_logits=True)
        # full-sequence causal decoding in transformers == our step-by-step decoding with the static cache
        ids = torch.tensor([o.sot_sequence("zh") + r["tokens"][b][:5] for b in range(n_chunks)])
        hf_logits = hf(input_features=torch.from_numpy(mel), decoder_input_ids=ids).logits
    ours = np.stack(r["logits"])  # [6, B, V]: logits after consuming positions 3..8
    for i in range(6):
        diff = np.abs(hf_logits[:, 3 + i].numpy() - ours[i]).max()
        assert diff <= tol, "step %d: %g" % (i, diff)
    print("%s logits: %d steps within %.0e of transformers" % (arch, 6, tol))
