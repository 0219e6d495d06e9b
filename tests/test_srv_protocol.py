"""whisper_srv wire protocol without a GPU (--no-model: a stub answers instead of the model): routes, status codes and
bodies of /root/reference/cpp/src/WhisperHTTPServer.hpp:39-117, HTTP/1.1 details (keep-alive, Expect: 100-continue, chunked
bodies), command line of /root/reference/cpp/whisper_srv.cpp:10-70."""
import http.client
import json
import os
import socket
import subprocess

import numpy as np
import pytest

import srv_util


@pytest.fixture(scope="module")
def srv(tmp_path_factory):
    if not os.path.exists(srv_util.SRV):
        pytest.fail("whisper_srv is not built (python -c 'import __graft_entry__ as g; g.build()')")
    s = srv_util.Server(["--no-model", "-t", "micro", "-p", "/nonexistent", "-l", "zh"], tmp_path_factory.mktemp("srv"))
    yield s
    assert s.stop() == 0


def test_asr_success_shape_and_cors(srv):
    st, h, body = srv.post_pcm(np.zeros(16000, np.float32))
    assert st == 200 and h["content-type"] == "application/json"
    assert body.decode() == '{\n  "success": true,\n  "text": "stub: 16000 samples"\n}'  # nlohmann dump(2): sorted keys, 2 spaces
    assert json.loads(body) == {"success": True, "text": "stub: 16000 samples"}
    assert h["access-control-allow-origin"] == "*" and "POST" in h["access-control-allow-methods"]  # WhisperHTTPServer.hpp:112-117
    assert "X-Array-Size" in h["access-control-allow-headers"]


def test_asr_error_paths(srv):
    pcm = np.zeros(1000, np.float32).tobytes()
    st, h, body = srv.request("POST", "/asr", body=pcm, headers={"Content-Type": "text/plain"})
    assert st == 400 and json.loads(body) == {"error": "Content-Type must be application/octet-stream"}  # :49-54
    assert h["access-control-allow-origin"] == "*"  # CORS headers are set before the checks (:43-44)
    st, _, body = srv.request("POST", "/asr", body=pcm)  # no Content-Type at all
    assert st == 400 and json.loads(body)["error"].startswith("Content-Type")
    st, _, body = srv.request("POST", "/asr", body=b"", headers={"Content-Type": "application/octet-stream"})
    assert st == 400 and json.loads(body) == {"error": "Request body is empty"}  # :57-61
    st, _, body = srv.request("POST", "/asr", body=pcm[:-1], headers={"Content-Type": "application/octet-stream"})
    assert st == 400 and json.loads(body) == {"error": "Data size must be multiple of 4 bytes"}  # :64-70
    st, _, body = srv.post_pcm(np.zeros(100, np.float32))  # the model refuses < 201 samples -> "Run model failed!" (:76-81)
    assert st == 400 and json.loads(body) == {"error": "Run model failed!"}
    assert srv.request("GET", "/asr")[0] == 404 and srv.request("POST", "/other", body=b"x")[0] == 404 and srv.request("GET", "/")[0] == 404


def test_keep_alive_expect_continue_and_chunked(srv):
    c = http.client.HTTPConnection("127.0.0.1", srv.port, timeout=30)
    for n in (16000, 32000, 48000):  # three requests on one connection
        c.request("POST", "/asr", body=np.zeros(n, np.float32).tobytes(), headers={"Content-Type": "application/octet-stream; charset=binary"})
        r = c.getresponse()
        assert r.status == 200 and json.loads(r.read())["text"] == "stub: %d samples" % n
    c.close()
    # Expect: 100-continue (what curl sends for large bodies), raw socket
    body = np.zeros(4000, np.float32).tobytes()
    s = socket.create_connection(("127.0.0.1", srv.port), timeout=30)
    s.sendall(("POST /asr HTTP/1.1\r\nHost: x\r\nContent-Type: application/octet-stream\r\nContent-Length: %d\r\nExpect: 100-continue\r\n"
               "Connection: close\r\n\r\n" % len(body)).encode())
    first = s.recv(64)
    assert first.startswith(b"HTTP/1.1 100 Continue")
    s.sendall(body)
    data = b""
    while True:
        chunk = s.recv(65536)
        if not chunk:
            break
        data += chunk
    assert b"200 OK" in (first + data) and b"stub: 4000 samples" in data
    s.close()
    # chunked request body
    s = socket.create_connection(("127.0.0.1", srv.port), timeout=30)
    s.sendall(b"POST /asr HTTP/1.1\r\nHost: x\r\nContent-Type: application/octet-stream\r\nTransfer-Encoding: chunked\r\nConnection: close\r\n\r\n")
    for part in (body[:6000], body[6000:]):
        s.sendall(b"%x\r\n" % len(part) + part + b"\r\n")
    s.sendall(b"0\r\n\r\n")
    data = b""
    while True:
        chunk = s.recv(65536)
        if not chunk:
            break
        data += chunk
    assert b"200 OK" in data and b"stub: 4000 samples" in data
    s.close()
    # POST without a length
    s = socket.create_connection(("127.0.0.1", srv.port), timeout=30)
    s.sendall(b"POST /asr HTTP/1.1\r\nHost: x\r\nContent-Type: application/octet-stream\r\n\r\n")
    assert b"411" in s.recv(4096)
    s.close()


def test_stats_and_cli(srv):
    st, _, body = srv.request("GET", "/stats")
    assert st == 200 and set(json.loads(body)) == {"requests", "gpu_passes"}
    out = subprocess.run([srv_util.SRV, "--help"], capture_output=True, text=True)
    assert out.returncode == 0 and "--model_type" in out.stderr and "--port" in out.stderr
    assert subprocess.run([srv_util.SRV, "--bogus"], capture_output=True).returncode != 0
    # without --no-model and without a model directory / GPU the server refuses to start (no CPU fallback)
    out = subprocess.run([srv_util.SRV, "--port", "0", "-t", "micro", "-p", "/nonexistent"], capture_output=True, text=True, timeout=120)
    assert out.returncode != 0 and "init server failed!" in out.stdout and "port: 0" in out.stdout and "model_type: micro" in out.stdout
