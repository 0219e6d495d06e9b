#!/usr/bin/env python3
"""bench.py -- headline benchmark: audio-seconds transcribed per second, Whisper-small, 30 s chunks, greedy 224 tokens.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (PCM -> log-mel -> encoder + cross K/V -> 4 SOT + 224 greedy decoder steps ->
token ids) over one batch of synthetic 30 s chunks per GPU (BASELINE.json configs[2]: Whisper-small, batch 256).
Shards are independent (no data-path collective): every rank owns its own chunks and weights;
torch.distributed (NCCL) is used only for the barrier and the max-over-ranks of the device time.

JSON line (rank 0): `value` = whole-job audio-s/s with PCM already resident in HBM (CUDA-event time, max over ranks);
`e2e` = same through the C ABI with pinned HOST buffers (H2D of the PCM and D2H of the token ids inside the timed
region); `roofline` = the dominant kernel (cross-attention decode, HBM-bound) timed alone with CUDA events on the
engine's stream; `cpu_baseline` = the CPU oracle (reference mel frontend + fp32 restatement of the exported graphs) on a
bounded sample of the same workload.  `--impl reference` times that CPU path as its own arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ARCH_DEFAULT = "small"
BATCH_DEFAULT = 256
NEW_TOKENS = 224
CHUNK_S = 30.0
CHUNK_SAMPLES = 480000


def shard_range(n_items, rank, world):
    """Static contiguous split of a list of independent utterances / 30 s windows across ranks:
    rank r owns [lo, hi); sizes differ by at most one; concatenating the ranks' results in rank order restores the order."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


def synth_batch(batch, rank, pinned=True):
    """Distribution N of SURVEY.md 8(d): 0.1*N(0,1) clipped to [-1,1]; seed = 1000*config + global chunk index."""
    import torch

    t = torch.empty((batch, CHUNK_SAMPLES), dtype=torch.float32, pin_memory=pinned and torch.cuda.is_available())
    a = t.numpy()
    for i in range(batch):
        rng = np.random.default_rng(2000 + rank * batch + i)
        a[i] = np.clip(0.1 * rng.standard_normal(CHUNK_SAMPLES, dtype=np.float32), -1, 1)
    return t, a


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (profiling recipe)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        # only samples taken under load (clock above idle) describe the timed region
        load = [s for s in sm if s > 500] or sm
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_pass(arch, n_chunks, new_tokens, seed_base):
    """The reference's CPU path on this box's host cores: its own C++ log-mel frontend (oracle/_ref, compiled from the
    reference sources) when present (else the numpy port) + the fp32 torch restatement of its exported encoder/decoder
    graphs (onnxruntime is not installed in this image).  Returns seconds."""
    import torch

    import util

    oracle = util.load_oracle(arch)
    audios = [np.clip(0.1 * np.random.default_rng(seed_base + i).standard_normal(CHUNK_SAMPLES, dtype=np.float32), -1, 1) for i in range(n_chunks)]
    t0 = time.perf_counter()
    mel = util.reference_mel(audios, oracle.n_mels)
    with torch.no_grad():
        oracle.transcribe_tokens(mel, max_new_tokens=new_tokens, honor_eot=False)
    return time.perf_counter() - t0


def run_reference_arm(args, rank, world):
    """--impl reference: the CPU path timed as its own arm (rank 0 only; other ranks exit without work)."""
    if rank != 0:
        return
    import torch

    import util

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_chunks = 2
    util.load_oracle(args.arch)
    for w in range(args.warmup):
        cpu_reference_pass(args.arch, n_chunks, 8, 9000 + w)  # short warm-up passes (page in weights, thread pools)
    times = [cpu_reference_pass(args.arch, n_chunks, args.new_tokens, 9100 + s) for s in range(args.steps)]
    dt = float(np.mean(times))
    value = CHUNK_S * n_chunks / dt
    kind = "reference+port" if util.mel_ref_lib() is not None else "port"
    line = {
        "impl": "reference", "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {**workload_config(args, world), "reference_arm": "each step is a bounded sample of this workload: %d x 30 s chunk(s) on the host CPU" % n_chunks},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port" if kind == "port" else "reference",
                         "sample": "%d x 30 s chunk per step, Whisper-%s, 4 SOT + %d greedy steps; mel = %s, encoder/decoder = fp32 torch "
                                   "restatement of the exported graphs (onnxruntime absent)" % (n_chunks, args.arch, args.new_tokens,
                                   "reference C++ frontend (oracle/_ref)" if kind != "port" else "numpy port")},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world, per_step_chunks=None):
    return {"workload": "Whisper-%s random-init, %d x 30 s synthetic 16 kHz chunks per GPU, 4 SOT + %d greedy decoder steps (EOT ignored), "
                        "zh" % (args.arch, args.batch if per_step_chunks is None else per_step_chunks, args.new_tokens),
            "arch": args.arch, "batch_per_gpu": args.batch if per_step_chunks is None else per_step_chunks,
            "global_batch": (args.batch if per_step_chunks is None else per_step_chunks) * (world if per_step_chunks is None else 1),
            "new_tokens": args.new_tokens, "parallelism": "dp%d (independent shards, no collective)" % world}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--arch", default=ARCH_DEFAULT)
    ap.add_argument("--batch", type=int, default=BATCH_DEFAULT, help="30 s chunks per GPU per step")
    ap.add_argument("--new-tokens", type=int, default=NEW_TOKENS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import __graft_entry__ as g
    import util

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL's own banner / warnings (NCCL_DEBUG=VERSION|WARN print to stdout) go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pkg = g.load_package()
    if rank == 0:
        util.model_root(args.arch)  # one rank generates the seeded random-init model directory
    barrier()
    eng = pkg.Engine(util.model_root(args.arch), args.arch, device=local_rank, max_batch=args.batch)
    B = args.batch
    pcm_t, pcm = synth_batch(B, rank)
    dims = eng.dims
    d, H, L = dims.d_model, dims.n_head, dims.n_text_layer

    # ---- `value`: inputs resident in HBM --------------------------------------------------------------------------
    eng.upload_pcm(pcm)
    for _ in range(args.warmup):
        eng.transcribe_resident(B, max_new_tokens=args.new_tokens, honor_eot=False)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    dev_ms, stage, launches = [], [], 0
    torch.cuda.nvtx.range_push("timed")  # `ncu --nvtx --nvtx-include "timed/"` profiles exactly the timed steps
    for _ in range(args.steps):
        toks, t = eng.transcribe_resident(B, max_new_tokens=args.new_tokens, honor_eot=False)
        dev_ms.append(t["total_ms"])
        stage.append(t)
        launches += t["kernel_launches"]
    torch.cuda.nvtx.range_pop()
    barrier()
    wall_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    assert all(len(x) == args.new_tokens for x in toks), "decode did not produce the configured number of tokens"

    # ---- `e2e`: through the C ABI with pinned host buffers ----------------------------------------------------------
    for _ in range(2):
        eng.transcribe(pcm, max_new_tokens=args.new_tokens, honor_eot=False)
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        eng.transcribe(pcm, max_new_tokens=args.new_tokens, honor_eot=False)
    barrier()
    e2e_wall = time.perf_counter() - t1

    # ---- dominant kernel alone (cross-attention decode: one launch per decoder layer) ----------------------------------
    eng.time_stage(3, B, iters=2)
    n_it = 8
    xattn_ms = min(eng.time_stage(3, B, iters=n_it) for _ in range(3)) / (n_it * L)  # best of 3 x (8 x L) launches
    mel_ms = float(np.mean([s["mel_ms"] for s in stage]))
    enc_ms = float(np.mean([s["encoder_ms"] for s in stage]))
    dec_ms = float(np.mean([s["decode_ms"] for s in stage]))

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    step_ms = allmax(float(np.mean(dev_ms)))
    wall_step_s = allmax(wall_s / args.steps)
    e2e_step_s = allmax(e2e_wall / args.steps)
    xattn_ms = allmax(xattn_ms)

    if rank == 0:
        pk = peaks()
        audio_s = CHUNK_S * B * world
        value = audio_s / (step_ms / 1e3)
        # algorithmic bytes of one cross-attention launch: K and V of every (sequence, head), bf16
        xattn_bytes = B * 1500 * d * 2 * 2
        achieved = xattn_bytes / (xattn_ms / 1e3) / 1e9
        # large batches run the streaming kernel (csrc/decode_ops.cu: launch_cross_attention_decode)
        xattn_kernel = "cross_attention_stream_kernel" if B * dims.n_head >= 296 else "cross_attention_decode_kernel"
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(xattn_kernel, {}).get("%s_b%d" % (args.arch, B))
        steps_dec = 4 + args.new_tokens
        w_dec = L * 14 * d * d * 2 + dims.n_vocab * d * 2
        dec_bytes = steps_dec * (w_dec + B * (L * 2 * 1500 * d * 2)) + B * L * 2 * d * 2 * steps_dec * (steps_dec + 1) // 2
        enc_flops = {"tiny": 40.48e9, "base": 96.80e9, "small": 386.63e9, "turbo": 2313.09e9}.get(args.arch)
        line = {
            "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {**workload_config(args, world), "l2": "inputs larger than L2 (cross K/V cache %.1f GB per GPU vs 126 MB L2)"
                       % (2 * L * B * 1500 * d * 2 / 1e9)},
            "clocks": clocks,
            "e2e": {"value": audio_s / e2e_step_s, "unit": "audio-s/s", "h2d_bytes_per_step": B * CHUNK_SAMPLES * 4 + B * 4,
                    "d2h_bytes_per_step": B * 448 * 4, "timing": "wall clock around the C-ABI call, barrier + synchronize both sides"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": xattn_kernel, "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk["source"] + " (burst copy figure; kernel timed alone; a read-only stream can exceed the read+write copy rate)",
                         "bytes_per_launch": xattn_bytes, "ms_per_launch": xattn_ms},
            "stages": {"mel_ms": mel_ms, "encoder_ms": enc_ms, "decode_ms": dec_ms, "wall_ms_per_step": wall_step_s * 1e3,
                       "mel_frac_hbm": (B * (1920000 + dims.n_mels * 3000 * 4) / (mel_ms / 1e3) / 1e9) / pk["hbm_gbs"] if mel_ms > 0 else None,
                       "encoder_frac_tensor_sustained": (enc_flops * B / (enc_ms / 1e3) / 1e12) / pk["bf16_tflops_sustained"] if enc_flops else None,
                       "decode_frac_hbm": (dec_bytes / (dec_ms / 1e3) / 1e9) / pk["hbm_gbs"]},
        }
        if not args.no_cpu_baseline and world == 1:
            import torch as _t

            cores = os.cpu_count() or 1
            _t.set_num_threads(cores)
            n_chunks = 6  # ~15 s of CPU work on this box's 16 host cores
            sec = cpu_reference_pass(args.arch, n_chunks, args.new_tokens, 7000)
            line["cpu_baseline"] = {
                "value": CHUNK_S * n_chunks / sec, "unit": "audio-s/s", "cores": cores,
                "kind": "reference" if util.mel_ref_lib() is not None else "port",
                "sample": "%d x 30 s chunks, Whisper-%s, 4 SOT + %d greedy steps, one pass (%.1f s): mel = %s; encoder/decoder = fp32 torch "
                          "restatement of the reference's exported graphs (onnxruntime not installed)"
                          % (n_chunks, args.arch, args.new_tokens, sec,
                             "the reference's own C++ frontend (oracle/_ref)" if util.mel_ref_lib() is not None else "numpy port")}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
