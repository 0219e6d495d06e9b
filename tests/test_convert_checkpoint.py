"""tools/convert_checkpoint.py: a transformers Whisper (random-init here: no checkpoints in this image) converted to the
engine's model directory gives the oracle the same logits as transformers computes -- i.e. the name mapping, the two-file
split and the derived configuration are right for checkpoints that were not produced by tools/make_model.py."""
import numpy as np
import pytest
import torch

import make_model
import util

transformers = pytest.importorskip("transformers")


def test_hf_checkpoint_round_trip(tmp_path):
    import convert_checkpoint
    import whisper_oracle

    torch.manual_seed(11)
    cfg = transformers.WhisperConfig(
        vocab_size=51865, num_mel_bins=80, d_model=128, encoder_layers=2, decoder_layers=3, encoder_attention_heads=2,
        decoder_attention_heads=2, encoder_ffn_dim=512, decoder_ffn_dim=512, max_source_positions=1500, max_target_positions=448,
        activation_function="gelu", dropout=0.0, attention_dropout=0.0, activation_dropout=0.0, scale_embedding=False)
    hf = transformers.WhisperForConditionalGeneration(cfg).eval()
    W = convert_checkpoint.from_hf_state_dict(hf.state_dict())
    dims = convert_checkpoint.check(W)
    assert dims == dict(n_mels=80, d=128, heads=2, l_enc=2, l_dec=3, n_vocab=51865)
    make_model.build_model_dir(str(tmp_path), "mine", weights=W)
    W2, c2 = make_model.load_model_dir(str(tmp_path), "mine")
    assert (c2["n_text_layer"], c2["n_audio_layer"], c2["n_text_state"], c2["n_mels"]) == (3, 2, 128, 80)
    assert set(W2) == set(W) and all(np.array_equal(W2[k], W[k]) for k in W)
    o = whisper_oracle.Oracle(W2, c2)
    mel = (np.random.default_rng(5).random((1, 80, 3000), dtype=np.float32) * 2 - 1)
    with torch.no_grad():
        ck, cv = o.encoder(mel)
        r = o.greedy(ck, cv, max_new_tokens=4, honor_eot=False, keep_logits=True)
        ids = torch.tensor([o.sot_sequence("zh") + r["tokens"][0][:3]])
        hf_logits = hf(input_features=torch.from_numpy(mel), decoder_input_ids=ids).logits
    for i in range(4):
        assert np.abs(hf_logits[:, 3 + i].numpy() - np.stack(r["logits"])[i]).max() <= 2e-4


def test_known_model_name_does_not_override_the_checkpoint_shape(tmp_path):
    """converting under one of the reference's model_type names (tiny / base / small / turbo) must still derive the
    configuration from the tensors -- a deeper checkpoint saved as "turbo" would otherwise get turbo's 4-decoder-layer table."""
    W = make_model.init_weights("micro")  # d = 128, 2 + 2 layers: not what the name "tiny" stands for
    make_model.build_model_dir(str(tmp_path), "tiny", weights=W)
    _, cfg = make_model.load_model_dir(str(tmp_path), "tiny")
    a = make_model.ARCHS["micro"]
    assert (cfg["n_text_state"], cfg["n_audio_layer"], cfg["n_text_layer"], cfg["n_text_head"]) == (a["d"], a["l_enc"], a["l_dec"], a["heads"])
    assert cfg["n_text_state"] != make_model.ARCHS["tiny"]["d"]
