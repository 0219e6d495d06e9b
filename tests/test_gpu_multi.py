"""Library-level multi-GPU sharding: a handle created with B200W_DEVICES=all splits a batch into
contiguous shards, one host thread and one engine per GPU, no collective.  Needs >= 2 visible GPUs (skipped otherwise)."""
import os

import pytest
import torch

import util

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_batch_sharded_over_all_gpus_matches_single_gpu(pkg):
    audios = [util.synth_audio("NUS"[i % 3], 160000 + 8000 * i, 70 + i) for i in range(7)]
    one = pkg.Whisper("micro", util.model_root("micro"), "zh")
    ref = one.run_tokens(audios, max_new_tokens=24, honor_eot=False)
    one.close()
    os.environ["B200W_DEVICES"] = "all"
    try:
        many = pkg.Whisper("micro", util.model_root("micro"), "zh")
        got = many.run_tokens(audios, max_new_tokens=24, honor_eot=False)
        long_audio = util.synth_audio("S", 480000 * 3 + 100000, 5)
        text_many = many.run_long(long_audio)
        many.close()
    finally:
        del os.environ["B200W_DEVICES"]
    assert got == ref
    one = pkg.Whisper("micro", util.model_root("micro"), "zh")
    assert one.run_long(long_audio) == text_many
    one.close()
