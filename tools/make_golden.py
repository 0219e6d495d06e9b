#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ (run in the build container, where /root/reference exists).

  mel_golden.npz      outputs of the reference's OWN C++ frontend (oracle/_ref/libmel_ref.so, compiled from
                      /root/reference/cpp/src) on seeded synthetic audio: one demo.wav-shaped clip in full and
                      strided samples of 30 s chunks for the N / U / S distributions at 80 and 128 mel bins.
  tiktoken_edge_cases.json  the awkward entries of the reference's BPE asset (NUL token, 25..33-byte tokens, UTF-8 fragments)
  oracle_micro.npz    outputs of the torch-CPU restatement (oracle/whisper_oracle.py) for the `micro` and
  oracle_tiny.npz     `tiny` architectures on a seeded mel input: strided cross K/V, per-step top-8 logits,
                      top-2 margins and greedy tokens.  The reference has no runnable encoder/decoder here
                      (parity unpinned), so these pin the oracle against regressions and give
                      the GPU box something to compare with that does not depend on re-running torch.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
FRAME_STRIDE = 25


def main():
    import torch

    os.makedirs(OUT, exist_ok=True)
    assert util.mel_ref_lib() is not None, "build oracle/_ref first (make -C oracle)"
    g = {}
    a = util.synth_audio("S", 67263, 1)
    g["short_S_80"] = util.reference_mel([a], 80)[0][:, :430]
    for dist in "NUS":
        for n_mels in (80, 128):
            a = util.synth_audio(dist, 480000, seed=100 + ord(dist))
            g["full_%s_%d" % (dist, n_mels)] = util.reference_mel([a], n_mels)[0][:, ::FRAME_STRIDE]
    np.savez_compressed(os.path.join(OUT, "mel_golden.npz"), **g)
    for arch in ("micro", "tiny"):
        o = util.load_oracle(arch)
        audios = [util.synth_audio("S", 480000, 7), util.synth_audio("N", 200000, 8)]
        mel = util.reference_mel(audios, o.n_mels)
        with torch.no_grad():
            ck, cv = o.encoder(mel)
            r = o.greedy(ck, cv, max_new_tokens=24, honor_eot=False, keep_logits=True)
        logits = np.stack(r["logits"])
        top_idx = np.argsort(-logits, axis=-1)[..., :8].astype(np.int32)
        top_val = np.take_along_axis(logits, top_idx, -1).astype(np.float32)
        np.savez_compressed(os.path.join(OUT, "oracle_%s.npz" % arch),
                            cross_k=ck.numpy()[:, :, ::50, ::4].astype(np.float16), cross_v=cv.numpy()[:, :, ::50, ::4].astype(np.float16),
                            tokens=np.array(r["tokens"], np.int32), margins=np.stack(r["top2_margin"]).astype(np.float32),
                            top_idx=top_idx, top_val=top_val)
    # edge cases of the reference's BPE vocabulary (python/assets/multilingual.tiktoken; /root/reference is absent on the GPU
    # box): the NUL token, every token longer than 24 bytes (the reference detokenises into char[32]), a few UTF-8 fragments
    import base64
    import json
    asset = "/root/reference/python/assets/multilingual.tiktoken"
    rows = [l.split() for l in open(asset) if l.strip()]
    pick = {}
    for tok, rank in rows:
        raw = base64.b64decode(tok)
        if b"\x00" in raw or len(raw) > 24 or int(rank) in (220, 11, 13, 50256) or (len(pick) < 400 and raw[:1] >= b"\xe0" and int(rank) % 37 == 0):
            pick[rank] = tok
    json.dump({"source": "multilingual.tiktoken of the reference (python/assets), ids -> base64 token bytes", "tokens": pick},
              open(os.path.join(OUT, "tiktoken_edge_cases.json"), "w"), indent=0)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
