"""CPU oracle for the encoder / decoder graphs and the greedy loop.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module -- as the checker (or the timed CPU baseline), never as part of the product path.

PARITY UNPINNED by the reference's own tests: /root/reference holds no golden tensors or
known-answer tests for this path and its execution stack (onnx,
onnxruntime, openai-whisper, the AXera NPU runtime) is absent from this image, so this is a
line-by-line fp32 torch-CPU restatement of what the reference exports and runs:

  * graphs        /root/reference/model_convert/export_onnx.py
                  encoder wrapper :153-213, self-attention with static cache :103-147,
                  cross/self wrappers :216-261, block :264-299, decoder step :302-387, mask :59-68
  * building blocks are openai-whisper==20240930 whisper/model.py (pinned in
                  /root/reference/model_convert/requirements.txt:1; NOT vendored in the reference):
                  LayerNorm(eps 1e-5, fp32), Linear, Conv1d, sinusoids, MultiHeadAttention
                  (key has no bias; scale (d/H)^-0.25 applied to q and k; softmax in fp32),
                  ResidualAttentionBlock (pre-LN; MLP = Linear, exact-erf GELU, Linear),
                  AudioEncoder, TextDecoder (logits tied to token_embedding).  Restated from the
                  published model definition; see SURVEY.md App. A.2-A.4.
  * greedy loop   /root/reference/cpp/src/Whisper.cpp:186-222 (run) and :290-346 (run_decoder):
                  4 SOT steps, then generate while idx != eot and offset < n_text_ctx; the host
                  writes this_self_k/v into cache row `offset` after each step (:328-342); argmax is
                  first-max (:42-45).
It is cross-checked against an independent implementation (transformers' Whisper, random init) in
tests/test_oracle_vs_transformers.py, and its outputs on seeded inputs are committed under
tests/golden/ by tools/make_golden.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

N_AUDIO_CTX = 1500
N_TEXT_CTX = 448
MASK_VALUE = -60000.0  # export_onnx.py:130 (not -inf)


def sinusoids(length, channels, max_timescale=10000.0):
    """whisper/model.py sinusoids(): [length, channels] = cat(sin, cos)."""
    assert channels % 2 == 0
    log_timescale_increment = math.log(max_timescale) / (channels // 2 - 1)
    inv_timescales = torch.exp(-log_timescale_increment * torch.arange(channels // 2, dtype=torch.float32))
    scaled_time = torch.arange(length, dtype=torch.float32)[:, None] * inv_timescales[None, :]
    return torch.cat([torch.sin(scaled_time), torch.cos(scaled_time)], dim=1)


class Oracle:
    """fp32 restatement. `weights` is {state_dict name: ndarray/tensor}; `cfg` the reference's _config.json dict."""

    def __init__(self, weights, cfg):
        self.W = {k: torch.from_numpy(np.array(v, dtype=np.float32)) for k, v in weights.items()}
        self.cfg = cfg
        self.d = int(cfg["n_text_state"])
        self.n_head_audio = int(cfg["n_audio_head"])
        self.n_head_text = int(cfg["n_text_head"])
        self.l_enc = int(cfg["n_audio_layer"])
        self.l_dec = int(cfg["n_text_layer"])
        self.n_vocab = int(cfg["n_vocab"])
        self.n_mels = int(cfg["n_mels"])
        self.pos_audio = sinusoids(N_AUDIO_CTX, int(cfg["n_audio_state"]))

    # ---- building blocks (whisper/model.py) -------------------------------------------------
    def _ln(self, x, prefix):
        return F.layer_norm(x.float(), (x.shape[-1],), self.W[prefix + ".weight"], self.W[prefix + ".bias"], 1e-5)

    def _lin(self, x, prefix, bias=True):
        return F.linear(x, self.W[prefix + ".weight"], self.W[prefix + ".bias"] if bias else None)

    @staticmethod
    def _qkv_attention(q, k, v, n_head):
        """MultiHeadAttention.qkv_attention with SDPA disabled (export_onnx.py:714-717), no mask."""
        n_batch, n_ctx, n_state = q.shape
        scale = (n_state // n_head) ** -0.25
        q = q.view(*q.shape[:2], n_head, -1).permute(0, 2, 1, 3)
        k = k.view(*k.shape[:2], n_head, -1).permute(0, 2, 1, 3)
        v = v.view(*v.shape[:2], n_head, -1).permute(0, 2, 1, 3)
        qk = (q * scale) @ (k * scale).transpose(-1, -2)
        w = F.softmax(qk.float(), dim=-1).to(q.dtype)
        return (w @ v).permute(0, 2, 1, 3).flatten(start_dim=2)

    def _mlp(self, x, prefix):
        return self._lin(F.gelu(self._lin(x, prefix + ".0")), prefix + ".2")

    # ---- encoder graph: export_onnx.py:153-213 --------------------------------------------------
    def conv_stem(self, mel):
        x = F.gelu(F.conv1d(mel, self.W["encoder.conv1.weight"], self.W["encoder.conv1.bias"], padding=1))
        x = F.gelu(F.conv1d(x, self.W["encoder.conv2.weight"], self.W["encoder.conv2.bias"], stride=2, padding=1))
        x = x.permute(0, 2, 1)
        return x + self.pos_audio[: x.shape[1]]

    def encoder_block(self, x, i):
        p = "encoder.blocks.%d" % i
        h = self._ln(x, p + ".attn_ln")
        q = self._lin(h, p + ".attn.query")
        k = self._lin(h, p + ".attn.key", bias=False)
        v = self._lin(h, p + ".attn.value")
        x = x + self._lin(self._qkv_attention(q, k, v, self.n_head_audio), p + ".attn.out")
        x = x + self._mlp(self._ln(x, p + ".mlp_ln"), p + ".mlp")
        return x

    def audio_features(self, mel, return_layers=False):
        """mel [B, n_mels, 3000] -> ln_post(x) [B, 1500, d]."""
        x = self.conv_stem(torch.as_tensor(mel).float())
        layers = [x]
        for i in range(self.l_enc):
            x = self.encoder_block(x, i)
            layers.append(x)
        x = self._ln(x, "encoder.ln_post")
        return (x, layers) if return_layers else x

    def encoder(self, mel):
        """-> cross_k, cross_v  [L_dec, B, 1500, d]  (AudioEncoderTensorCache.forward, :193-213)."""
        xa = self.audio_features(mel)
        ks, vs = [], []
        for i in range(self.l_dec):
            p = "decoder.blocks.%d.cross_attn" % i
            ks.append(self._lin(xa, p + ".key", bias=False))
            vs.append(self._lin(xa, p + ".value"))
        return torch.stack(ks, 0), torch.stack(vs, 0)

    # ---- decoder step: export_onnx.py:302-387 ---------------------------------------------------
    def decoder_step(self, tokens, self_k, self_v, cross_k, cross_v, offset, mask):
        """tokens [B] int; self_k/v [L,B,448,d]; cross_k/v [L,B,1500,d]; offset int; mask [448] (1 = masked).
        Returns logits [B,V], this_self_k [L,B,d], this_self_v [L,B,d].  (The reference asserts B == 1,
        :333; the arithmetic is per-sequence, so batching the oracle changes nothing.)"""
        W = self.W
        tokens = torch.as_tensor(tokens).long().view(-1)
        B = tokens.shape[0]
        H = self.n_head_text
        x = (W["decoder.token_embedding.weight"][tokens] + W["decoder.positional_embedding"][int(offset)]).unsqueeze(1)
        mask_b = torch.as_tensor(mask).bool()
        ks, vs = [], []
        for i in range(self.l_dec):
            p = "decoder.blocks.%d" % i
            # self attention over the static 448-slot cache plus the current token (:103-147, :233-261)
            h = self._ln(x, p + ".attn_ln")
            q = self._lin(h, p + ".attn.query")
            k1 = self._lin(h, p + ".attn.key", bias=False)
            v1 = self._lin(h, p + ".attn.value")
            scale = (self.d // H) ** -0.25
            qh = q.view(B, 1, H, -1).permute(0, 2, 1, 3)
            kc = self_k[i].view(B, N_TEXT_CTX, H, -1).permute(0, 2, 1, 3)
            vc = self_v[i].view(B, N_TEXT_CTX, H, -1).permute(0, 2, 1, 3)
            k1h = k1.view(B, 1, H, -1).permute(0, 2, 1, 3)
            v1h = v1.view(B, 1, H, -1).permute(0, 2, 1, 3)
            qk = (qh * scale) @ (kc * scale).transpose(-1, -2)
            qk1 = (qh * scale) @ (k1h * scale).transpose(-1, -2)
            qk = qk.masked_fill(mask_b, MASK_VALUE)
            w_total = F.softmax(torch.cat([qk.float(), qk1.float()], dim=-1), dim=-1)
            out = (w_total[..., :-1] @ vc).permute(0, 2, 1, 3).flatten(start_dim=2)
            out = out + (w_total[..., -1:] @ v1h).permute(0, 2, 1, 3).flatten(start_dim=2)
            x = x + self._lin(out, p + ".attn.out")
            # cross attention (:216-230), un-patched qkv_attention, no mask
            h = self._ln(x, p + ".cross_attn_ln")
            q = self._lin(h, p + ".cross_attn.query")
            x = x + self._lin(self._qkv_attention(q, cross_k[i], cross_v[i], H), p + ".cross_attn.out")
            x = x + self._mlp(self._ln(x, p + ".mlp_ln"), p + ".mlp")
            ks.append(k1[:, 0])
            vs.append(v1[:, 0])
        x = self._ln(x, "decoder.ln")
        logits = (x[:, 0] @ W["decoder.token_embedding.weight"].t()).float()
        return logits, torch.stack(ks, 0), torch.stack(vs, 0)

    # ---- greedy loop: Whisper.cpp:186-222, 290-346 ----------------------------------------------
    def sot_sequence(self, lang="zh"):
        codes = self.cfg["all_language_codes"].split(",")
        toks = [int(t) for t in self.cfg["all_language_tokens"].split(",")]
        if lang not in codes:
            lang = "zh"  # Whisper.cpp:244-248, DEFAULT_LANG
        return [int(self.cfg["sot"]), toks[codes.index(lang)], int(self.cfg["transcribe"]), int(self.cfg["no_timestamps"])]

    @torch.no_grad()
    def greedy(self, cross_k, cross_v, lang="zh", max_new_tokens=None, honor_eot=True, forced_tokens=None,
               keep_logits=False):
        """Batched greedy decode.  Returns dict(tokens=[B][...], logits=[steps][B,V] if keep_logits,
        top2_margin=[steps][B]).  forced_tokens [B, n] (optional): teacher forcing -- the i-th generated
        token fed back is forced_tokens[:, i] instead of the argmax (logits are still recorded)."""
        L, B = cross_k.shape[0], cross_k.shape[1]
        d = self.d
        eot = int(self.cfg["eot"])
        self_k = torch.zeros(L, B, N_TEXT_CTX, d)
        self_v = torch.zeros(L, B, N_TEXT_CTX, d)
        mask = torch.ones(N_TEXT_CTX, dtype=torch.int32)
        offset = 0
        out_tokens = [[] for _ in range(B)]
        done = [False] * B
        logits_log, margin_log = [], []

        def step(tok):
            nonlocal offset
            if offset > 0:
                mask[offset - 1] = 0  # causal_mask_1d, Whisper.cpp:253-258
            logits, k1, v1 = self.decoder_step(tok, self_k, self_v, cross_k, cross_v, offset, mask)
            self_k[:, :, offset] = k1  # Whisper.cpp:328-342
            self_v[:, :, offset] = v1
            offset += 1
            return logits

        sot = self.sot_sequence(lang)
        logits = None
        for t in sot:
            logits = step(torch.full((B,), t, dtype=torch.long))
        n_gen = 0
        limit = N_TEXT_CTX - len(sot) if max_new_tokens is None else max_new_tokens
        while True:
            top2 = torch.topk(logits, 2, dim=-1).values
            margin_log.append((top2[:, 0] - top2[:, 1]).numpy().copy())
            if keep_logits:
                logits_log.append(logits.numpy().copy())
            idx = torch.argmax(logits, dim=-1)  # first max, like std::max_element
            if forced_tokens is not None and n_gen < forced_tokens.shape[1]:
                feed = torch.as_tensor(forced_tokens[:, n_gen]).long()
            else:
                feed = idx
            for b in range(B):
                if done[b]:
                    continue
                if honor_eot and int(idx[b]) == eot:
                    done[b] = True
                else:
                    out_tokens[b].append(int(idx[b]))
            n_gen += 1
            if all(done) or n_gen >= limit or offset >= N_TEXT_CTX:
                break
            logits = step(feed)
        return dict(tokens=out_tokens, logits=logits_log, top2_margin=margin_log)

    @torch.no_grad()
    def transcribe_tokens(self, mel, **kw):
        ck, cv = self.encoder(mel)
        return self.greedy(ck, cv, **kw)
