"""whisper_srv on the GPU: the reference's HTTP front-end (/root/reference/cpp/src/WhisperHTTPServer.hpp:39-100) over
AX_WHISPER_RunPCM.  32 concurrent posts must be answered with the texts the serial calls give, and AX_WHISPER_GetStats
(GET /stats) must show that they were coalesced into fewer GPU passes than requests."""
import json
import threading

import numpy as np
import pytest

import srv_util
import util

pytestmark = pytest.mark.gpu


def test_concurrent_posts_are_coalesced_and_correct(pkg, tmp_path):
    root = util.model_root("tiny")
    audios = [util.synth_audio("NUS"[i % 3], 60000 + 9000 * i, 300 + i) for i in range(32)]
    w = pkg.Whisper("tiny", root, "zh")
    serial = [w.run(a) for a in audios[:8]] + [None] * 24
    toks = w.run_tokens(audios)
    w.close()
    detok = lambda t: "".join(" t%d" % x for x in t if x < 50257)  # synthetic token table of tools/make_model.py
    for i in range(8):
        assert serial[i] == detok(toks[i])
    srv = srv_util.Server(["-t", "tiny", "-p", root, "-l", "zh", "--coalesce_max", "32"], tmp_path)
    try:
        st, _, body = srv.post_pcm(audios[0])  # one request alone (also warms the engine up)
        assert st == 200 and json.loads(body) == {"success": True, "text": detok(toks[0])}
        before = json.loads(srv.request("GET", "/stats")[2])
        out = [None] * 32

        def post(i):
            out[i] = srv.post_pcm(audios[i], timeout=300)

        threads = [threading.Thread(target=post, args=(i,)) for i in range(32)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        after = json.loads(srv.request("GET", "/stats")[2])
        for i in range(32):
            st, h, body = out[i]
            assert st == 200 and h["access-control-allow-origin"] == "*"
            assert json.loads(body) == {"success": True, "text": detok(toks[i])}, "request %d" % i
        n_req, n_pass = after["requests"] - before["requests"], after["gpu_passes"] - before["gpu_passes"]
        print("32 concurrent posts were served in %d GPU passes" % n_pass)
        assert n_req == 32 and n_pass < 32, "no coalescing: %d requests in %d passes" % (n_req, n_pass)
        # error paths with the real model behind
        st, _, body = srv.post_pcm(np.zeros(100, np.float32))
        assert st == 400 and json.loads(body) == {"error": "Run model failed!"}
        st, _, body = srv.request("POST", "/asr", body=b"abc", headers={"Content-Type": "application/octet-stream"})
        assert st == 400 and json.loads(body) == {"error": "Data size must be multiple of 4 bytes"}
    finally:
        assert srv.stop() == 0
