"""bench.py's CPU legs (the only parts that run without a GPU): the reference arm's JSON contract and the configs[0] report."""
import json
import os
import subprocess
import sys

import util


def _run(args):
    out = subprocess.run([sys.executable, os.path.join(util.ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    return json.loads(lines[0])


def test_reference_arm_contract():
    d = _run(["--impl", "reference", "--arch", "tiny", "--steps", "1", "--warmup", "0", "--new-tokens", "6"])
    assert d["impl"] == "reference" and d["metric"] == "audio_seconds_per_second" and d["unit"] == "audio-s/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert "6 x 30 s chunks" in d["cpu_baseline"]["sample"]  # >= 6 chunks per step (2 under-state the CPU)
    assert d["e2e"] == {"value": d["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_config0_cpu_report():
    """BASELINE.json configs[0]: tiny, B = 1, per-stage ms, single-threaded reference mel, cores stated, both clip shapes."""
    d = _run(["--config", "0"])
    rep = d["config0"]
    assert d["impl"] == "reference" and d["config"]["config"] == 0 and rep["cores"] >= 1 and "1 thread" in rep["mel"]
    audios = {c["audio"] for c in rep["cases"]}
    assert audios == {"30 s chunk", "demo.wav shape, 67263 samples"}
    for c in rep["cases"]:
        assert c["decode_steps"] == 448 and c["mel_ms"] > 0 and c["encoder_ms"] > 0 and c["decode_ms"] > 0 and c["rtf"] > 0
        assert abs(c["total_ms"] - (c["mel_ms"] + c["encoder_ms"] + c["decode_ms"])) < 1.0


def test_gpu_arm_refuses_to_run_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(util.ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CPU path" in (out.stderr + out.stdout)
