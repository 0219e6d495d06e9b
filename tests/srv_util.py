"""Helpers for the whisper_srv tests: start the server executable on a free port, talk HTTP/1.1 to it with the standard library."""
import http.client
import os
import subprocess
import time

import util

SRV = os.path.join(util.ROOT, "whisper.axera_b200", "whisper_srv")


class Server:
    def __init__(self, args, tmp_path, env=None, start_timeout=300):
        self.port_file = str(tmp_path / "port.txt")
        self.log = open(str(tmp_path / "srv.log"), "w")
        e = dict(os.environ)
        e.update(env or {})
        self.proc = subprocess.Popen([SRV, "--port", "0", "--host", "127.0.0.1", "--port_file", self.port_file] + args, stdout=self.log,
                                     stderr=subprocess.STDOUT, env=e)
        t0 = time.time()
        while not os.path.exists(self.port_file):
            if self.proc.poll() is not None:
                raise RuntimeError("whisper_srv exited with %s: %s" % (self.proc.returncode, open(str(tmp_path / "srv.log")).read()))
            if time.time() - t0 > start_timeout:
                self.proc.kill()
                raise RuntimeError("whisper_srv did not start")
            time.sleep(0.05)
        self.port = int(open(self.port_file).read())

    def request(self, method, path, body=None, headers=None, timeout=120):
        c = http.client.HTTPConnection("127.0.0.1", self.port, timeout=timeout)
        c.request(method, path, body=body, headers=headers or {})
        r = c.getresponse()
        data = r.read()
        hdrs = {k.lower(): v for k, v in r.getheaders()}
        c.close()
        return r.status, hdrs, data

    def post_pcm(self, pcm, **kw):
        return self.request("POST", "/asr", body=pcm.astype("<f4").tobytes(), headers={"Content-Type": "application/octet-stream"}, **kw)

    def stop(self):
        if self.proc.poll() is None:
            self.proc.terminate()  # SIGTERM: graceful stop
            try:
                self.proc.wait(timeout=20)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        self.log.close()
        return self.proc.returncode
