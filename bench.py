#!/usr/bin/env python3
"""bench.py -- headline benchmark: audio-seconds transcribed per second, Whisper-small, 30 s chunks, greedy 224 tokens.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config {0,1,2,3,4}] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (PCM -> log-mel -> encoder + cross K/V -> 4 SOT + 224 greedy decoder steps ->
token ids) over one batch of synthetic 30 s chunks per GPU.  The default workload is BASELINE.json configs[2]
(Whisper-small, batch 256 per GPU); --config selects another BASELINE configuration:
    0  Whisper-tiny, B = 1, the reference's CPU-runnable case: CPU report only (per-stage ms, cores stated)
    1  Whisper-base, 64 x 30 s chunks on one GPU
    2  Whisper-small, 256 x 30 s chunks per GPU (default; the strong split 256 / N per GPU is reported under extra.strong)
    3  Whisper-turbo, 128 x 30 s chunks per GPU
    4  long-form: 1 h of audio = 120 x 30 s windows spread over the N GPUs, decoded until the 448-token context is full
Shards are independent (no data-path collective): every rank owns its own chunks and weights;
torch.distributed (NCCL) is used only for barriers, the max-over-ranks of the device time and gathering token ids for checks.

JSON line (rank 0): `value` = whole-job audio-s/s with PCM already resident in HBM (CUDA-event time, max over ranks);
`e2e` = the same through the reference-shaped C ABI (AX_WHISPER_RunPCMTokens) with pinned HOST buffers (H2D of the PCM and
D2H of the token ids inside the timed region); `roofline` = the dominant kernel (cross-attention decode, HBM-bound) timed alone
with CUDA events on the engine's stream, for the launch shape the step really issues (one micro-batch) and for the whole batch;
`cpu_baseline` = the CPU oracle (reference mel frontend + fp32 restatement of the exported graphs) on a bounded sample of the
same workload; `extra` = what the driver's weak-scaling curve cannot see: the strong split of configs[2], the library's own
multi-GPU path (one handle, B200W_DEVICES), and compact lines for configs 0, 1, 3, 4.  `--impl reference` times the CPU path
as its own arm.
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

NEW_TOKENS = 224
CHUNK_S = 30.0
CHUNK_SAMPLES = 480000
# BASELINE.json configs -> (architecture, chunks per GPU, generated tokens)
CONFIGS = {0: ("tiny", 1, 444), 1: ("base", 64, NEW_TOKENS), 2: ("small", 256, NEW_TOKENS), 3: ("turbo", 128, NEW_TOKENS), 4: ("small", 120, 444)}
ENC_FLOPS = {"tiny": 40.48e9, "base": 96.80e9, "small": 386.63e9, "turbo": 2313.09e9}  # SURVEY.md 8(d), per 30 s chunk


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


def synth_chunk(global_index):
    """Distribution N of SURVEY.md 8(d): 0.1*N(0,1) clipped to [-1,1]; seed = 2000 + global chunk index."""
    rng = np.random.default_rng(2000 + global_index)
    return np.clip(0.1 * rng.standard_normal(CHUNK_SAMPLES, dtype=np.float32), -1, 1)


def synth_batch(batch, first_index, pinned=True):
    import torch

    t = torch.empty((batch, CHUNK_SAMPLES), dtype=torch.float32, pin_memory=pinned and torch.cuda.is_available())
    a = t.numpy()
    for i in range(batch):
        a[i] = synth_chunk(first_index + i)
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


# ----------------------------------------------------------------------------------------------------------------------
# CPU legs (the only places that execute oracle/)
# ----------------------------------------------------------------------------------------------------------------------
def cpu_reference_pass(arch, n_chunks, new_tokens, seed_base, n_samples=CHUNK_SAMPLES, stages=None):
    """The reference's CPU path on this box's host cores: its own C++ log-mel frontend (oracle/_ref, compiled from the
    reference sources, single-threaded like the reference) when present (else the numpy port) + the fp32 torch restatement of
    its exported encoder/decoder graphs (onnxruntime is not installed in this image).  Returns seconds."""
    import torch

    import util

    oracle = util.load_oracle(arch)
    audios = [np.clip(0.1 * np.random.default_rng(seed_base + i).standard_normal(n_samples, dtype=np.float32), -1, 1) for i in range(n_chunks)]
    t0 = time.perf_counter()
    mel = util.reference_mel(audios, oracle.n_mels)
    t1 = time.perf_counter()
    with torch.no_grad():
        ck, cv = oracle.encoder(mel)
        t2 = time.perf_counter()
        r = oracle.greedy(ck, cv, max_new_tokens=new_tokens, honor_eot=False)
    t3 = time.perf_counter()
    if stages is not None:
        stages.update(mel_ms=(t1 - t0) * 1e3, encoder_ms=(t2 - t1) * 1e3, decode_ms=(t3 - t2) * 1e3, decode_steps=4 + len(r["tokens"][0]),
                      total_ms=(t3 - t0) * 1e3)
    return t3 - t0


def config0_report():
    """BASELINE.json configs[0] / BASELINE.md section 3: Whisper-tiny random-init, B = 1, greedy zh on the host CPU -- the repo's
    C++ log-mel (single thread, as the reference runs it) + the exported encoder / decoder graphs (fp32 torch restatement;
    onnxruntime absent) with 4 intra-op threads like the reference's generate_data.py:37-39 and with all cores; one 30 s chunk
    and one demo.wav-shaped clip (67 263 samples); per-stage ms and RTF."""
    import torch

    import util

    cores = os.cpu_count() or 1
    out = {"workload": "Whisper-tiny random-init, B=1, 4 SOT + 444 greedy decoder steps (random weights never emit EOT), zh", "cores": cores,
           "mel": "reference C++ frontend (oracle/_ref), 1 thread" if util.mel_ref_lib() is not None else "numpy port, 1 thread", "cases": []}
    util.load_oracle("tiny")
    cpu_reference_pass("tiny", 1, 4, 8000)  # page in weights / thread pools
    for threads in sorted({min(4, cores), cores}):
        torch.set_num_threads(threads)
        for name, n in (("30 s chunk", CHUNK_SAMPLES), ("demo.wav shape, 67263 samples", 67263)):
            st = {}
            sec = cpu_reference_pass("tiny", 1, 444, 8100, n_samples=n, stages=st)
            dur = n / 16000.0
            out["cases"].append({"audio": name, "torch_threads": threads, **{k: round(v, 2) if isinstance(v, float) else v for k, v in st.items()},
                                 "ms_per_decoder_step": round(st["decode_ms"] / st["decode_steps"], 3), "rtf": sec / dur, "audio_s_per_s": dur / sec})
    torch.set_num_threads(cores)
    return out


def run_reference_arm(args, rank, world):
    """--impl reference: the CPU path timed as its own arm (rank 0 only; other ranks exit without work)."""
    if rank != 0:
        return
    import torch

    import util

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if args.config == 0:
        rep = config0_report()
        best = max(c["audio_s_per_s"] for c in rep["cases"] if c["audio"].startswith("30 s"))
        print(json.dumps({"impl": "reference", "metric": "audio_seconds_per_second", "value": best, "unit": "audio-s/s", "n_gpus": args.gpus,
                          "steps": 1, "warmup": 1, "ms_per_step": CHUNK_S / best * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": rep["workload"], "config": 0},
                          "cpu_baseline": {"value": best, "unit": "audio-s/s", "cores": cores, "kind": "reference" if util.mel_ref_lib() is not None else "port",
                                           "sample": "one 30 s chunk, full 448-token context"},
                          "e2e": {"value": best, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
                          "config0": rep}), flush=True)
        return
    n_chunks = 6  # one encoder batch of 6 chunks keeps every host core busy (2 chunks under-state the CPU by ~40 %)
    util.load_oracle(args.arch)
    for w in range(args.warmup):
        cpu_reference_pass(args.arch, 2, 8, 9000 + w)  # short warm-up passes (page in weights, thread pools)
    times = [cpu_reference_pass(args.arch, n_chunks, args.new_tokens, 9100 + s) for s in range(args.steps)]
    dt = float(np.mean(times))
    value = CHUNK_S * n_chunks / dt
    kind = "reference" if util.mel_ref_lib() is not None else "port"
    line = {
        "impl": "reference", "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {**workload_config(args, world), "reference_arm": "each step is a bounded sample of this workload: %d x 30 s chunks on the host CPU" % n_chunks},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": kind,
                         "sample": "%d x 30 s chunks per step, Whisper-%s, 4 SOT + %d greedy steps; mel = %s, encoder/decoder = fp32 torch "
                                   "restatement of the exported graphs (onnxruntime absent)" % (n_chunks, args.arch, args.new_tokens,
                                   "reference C++ frontend (oracle/_ref)" if kind == "reference" else "numpy port")},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": "Whisper-%s random-init, %d x 30 s synthetic 16 kHz chunks per GPU, 4 SOT + %d greedy decoder steps (EOT ignored), zh"
                        % (args.arch, args.batch, args.new_tokens),
            "config": args.config, "arch": args.arch, "batch_per_gpu": args.batch, "global_batch": args.batch * world,
            "new_tokens": args.new_tokens, "parallelism": "dp%d (independent shards, no collective)" % world}


# ----------------------------------------------------------------------------------------------------------------------
# GPU legs
# ----------------------------------------------------------------------------------------------------------------------
class Dist:
    def __init__(self, rank, local_rank, world):
        self.rank, self.local_rank, self.world = rank, local_rank, world
        self.cpu_group = None
        if world > 1:
            import torch.distributed as dist

            try:
                self.cpu_group = dist.new_group(backend="gloo")  # host-side waits that leave the GPUs idle (an NCCL barrier spins on them)
            except Exception as ex:  # no usable interface for gloo: fall back to the NCCL barrier
                print("[bench] gloo group unavailable (%s): host waits use the NCCL barrier" % ex, file=sys.stderr, flush=True)

    def host_barrier(self):
        if self.world > 1:
            import torch.distributed as dist

            if self.cpu_group is not None:
                dist.barrier(group=self.cpu_group)
            else:
                dist.barrier()

    def barrier(self):
        import torch

        if self.world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def allmax(self, x):
        if self.world == 1:
            return float(x)
        import torch
        import torch.distributed as dist

        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_tokens(self, toks, n_cols):
        """[B][n_cols] token ids of every rank -> list over ranks on rank 0 (None elsewhere)."""
        import torch

        arr = np.array([t[:n_cols] + [-1] * (n_cols - len(t)) for t in toks], np.int32)
        if self.world == 1:
            return [arr]
        import torch.distributed as dist

        t = torch.from_numpy(arr).cuda()
        out = [torch.empty_like(t) for _ in range(self.world)] if self.rank == 0 else None
        dist.gather(t, out, dst=0)
        return [o.cpu().numpy() for o in out] if self.rank == 0 else None


def decode_bytes(dims, B, steps_dec):
    """Algorithmic HBM bytes of a greedy decode: per step W_dec + B * (cross K/V of every layer + the self K/V read so far)."""
    d, L = dims.d_model, dims.n_text_layer
    w_dec = L * 14 * d * d * 2 + dims.n_vocab * d * 2
    return steps_dec * (w_dec + B * (L * 2 * 1500 * d * 2)) + B * L * 2 * d * 2 * steps_dec * (steps_dec + 1) // 2


def resident_leg(eng, pcm, B, new_tokens, steps, warmup, dd, sampler=None):
    """`value` leg: PCM resident in HBM, CUDA-event times of whole passes; returns per-rank means + the last tokens."""
    import torch

    eng.upload_pcm(pcm[:B])
    for _ in range(warmup):
        eng.transcribe_resident(B, max_new_tokens=new_tokens, honor_eot=False)
    dd.barrier()
    if sampler is not None:
        sampler.start()
    t0 = time.perf_counter()
    dev_ms, stage, launches, toks = [], [], 0, None
    torch.cuda.nvtx.range_push("timed")  # `ncu --nvtx --nvtx-include "timed/"` profiles exactly the timed steps
    for _ in range(steps):
        toks, t = eng.transcribe_resident(B, max_new_tokens=new_tokens, honor_eot=False)
        dev_ms.append(t["total_ms"])
        stage.append(t)
        launches += t["kernel_launches"]
    torch.cuda.nvtx.range_pop()
    dd.barrier()
    wall_s = time.perf_counter() - t0
    assert all(len(x) == new_tokens for x in toks), "decode did not produce the configured number of tokens"
    mean = lambda k: float(np.mean([s[k] for s in stage]))
    return dict(step_ms=float(np.mean(dev_ms)), wall_step_s=wall_s / steps, launches=int(launches), toks=toks,
                mel_ms=mean("mel_ms"), enc_ms=mean("encoder_ms"), dec_ms=mean("decode_ms"))


def stage_fracs(r, dims, arch, B, new_tokens, pk):
    steps_dec = 4 + new_tokens
    fr = {"mel_ms": r["mel_ms"], "encoder_ms": r["enc_ms"], "decode_ms": r["dec_ms"], "ms_per_decoder_step": r["dec_ms"] / steps_dec,
          "mel_frac_hbm": (B * (1920000 + dims.n_mels * 3000 * 4) / (r["mel_ms"] / 1e3) / 1e9) / pk["hbm_gbs"] if r["mel_ms"] > 0 else None,
          "decode_frac_hbm": (decode_bytes(dims, B, steps_dec) / (r["dec_ms"] / 1e3) / 1e9) / pk["hbm_gbs"]}
    if arch in ENC_FLOPS and r["enc_ms"] > 0:
        tf = ENC_FLOPS[arch] * B / (r["enc_ms"] / 1e3) / 1e12
        fr["encoder_frac_tensor_sustained"] = tf / pk["bf16_tflops_sustained"]
        fr["encoder_frac_tensor_burst"] = tf / pk["bf16_tflops"]
    return fr


def compact_config_line(pkg, util, dd, arch, B, new_tokens, pk, steps=2, warmup=3, first_index=0):
    """One other BASELINE configuration, measured like the headline's `value` leg (resident PCM, CUDA events, max over ranks)."""
    if dd.rank == 0:
        util.model_root(arch)
    dd.barrier()
    eng = pkg.Engine(util.model_root(arch), arch, device=dd.local_rank, max_batch=B)
    try:
        _, pcm = synth_batch(B, first_index + dd.rank * B, pinned=False)
        r = resident_leg(eng, pcm, B, new_tokens, steps, warmup, dd)
        step_ms = dd.allmax(r["step_ms"])
        out = {"arch": arch, "batch_per_gpu": B, "n_gpus": dd.world, "new_tokens": new_tokens, "ms_per_step": step_ms,
               "audio_s_per_s": CHUNK_S * B * dd.world / (step_ms / 1e3), "stages": stage_fracs(r, eng.dims, arch, B, new_tokens, pk)}
    finally:
        eng.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configs[] index (default 2 = the headline)")
    ap.add_argument("--arch", default=None)
    ap.add_argument("--batch", type=int, default=None, help="30 s chunks per GPU per step")
    ap.add_argument("--new-tokens", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip extra.* (strong split, library path, other configs)")
    args = ap.parse_args()
    c_arch, c_batch, c_new = CONFIGS[args.config]
    args.arch = args.arch or c_arch
    args.new_tokens = args.new_tokens or c_new
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.batch is None:
        args.batch = c_batch if args.config != 4 else max(1, (c_batch + world - 1) // world)  # long-form: 120 windows over the GPUs
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference" or args.config == 0:
        if args.config == 0:
            args.impl = "reference"  # configs[0] is the reference's own CPU-runnable case: there is no GPU arm for it
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
    dd = Dist(rank, local_rank, world)
    t_start = time.perf_counter()

    def phase(name):  # wall-clock trace of the run on stderr (stdout carries exactly one JSON line)
        if rank == 0:
            print("[bench +%6.1f s] %s" % (time.perf_counter() - t_start, name), file=sys.stderr, flush=True)

    pk = peaks()
    pkg = g.load_package()
    if rank == 0:
        util.model_root(args.arch)  # one rank generates the seeded random-init model directory
    dd.barrier()
    B = args.batch
    eng = pkg.Engine(util.model_root(args.arch), args.arch, device=local_rank, max_batch=B)
    pcm_t, pcm = synth_batch(B, rank * B)
    dims = eng.dims
    d, H, L = dims.d_model, dims.n_head, dims.n_text_layer

    phase("engine ready, synthetic PCM generated")
    # ---- `value`: inputs resident in HBM --------------------------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    r = resident_leg(eng, pcm, B, args.new_tokens, args.steps, args.warmup, dd, sampler)
    clocks = sampler.stop() if rank == 0 else None
    toks_resident = r["toks"]
    step_ms = dd.allmax(r["step_ms"])
    wall_step_s = dd.allmax(r["wall_step_s"])

    # ---- dominant kernel alone (cross-attention decode, one launch per decoder layer): the launch shape the step issues
    # (a micro-batch of B/2 sequences when B >= 32) and the whole batch ------------------------------------------------
    def xattn_ms(nb):
        eng.time_stage(3, nb, iters=2)
        n_it = 8
        return dd.allmax(min(eng.time_stage(3, nb, iters=n_it) for _ in range(3)) / (n_it * L))  # best of 3 x (8 x L) launches

    phase("resident leg done")
    nb_step = (B + 1) // 2 if B >= 32 else B
    x_step_ms = xattn_ms(nb_step)
    x_full_ms = xattn_ms(B) if nb_step != B else x_step_ms

    # ---- strong split of this workload: the same global batch (B chunks in total) spread over the N GPUs --------------
    extra = {}
    strong = None
    if not args.no_extras and args.config == 2:
        lo, hi = shard_range(B, rank, world)
        if world == 1:
            strong = {"note": "N = 1: the strong split is the headline itself", "chunks_per_gpu": B, "ms_per_step": step_ms,
                      "audio_s_per_s": CHUNK_S * B / (step_ms / 1e3)}
        else:
            rs = resident_leg(eng, pcm, hi - lo, args.new_tokens, max(2, min(args.steps, 5)), 3, dd)
            s_ms = dd.allmax(rs["step_ms"])
            strong = {"global_batch": B, "chunks_per_gpu": hi - lo, "ms_per_step": s_ms, "audio_s_per_s": CHUNK_S * B / (s_ms / 1e3),
                      "scaling": "strong", "stages": stage_fracs(rs, dims, args.arch, hi - lo, args.new_tokens, pk)}
        extra["strong"] = strong
    eng.close()

    phase("roofline kernel + strong split done")
    # ---- `e2e`: the reference-shaped C ABI (AX_WHISPER_Init / AX_WHISPER_RunPCMTokens) with pinned host buffers ---------
    os.environ["B200W_DEVICE"] = str(local_rank)
    os.environ["B200W_MAX_BATCH"] = str(B)
    os.environ.pop("B200W_DEVICES", None)
    w = pkg.Whisper(args.arch, util.model_root(args.arch), "zh")
    rows = [pcm[i] for i in range(B)]
    for _ in range(2):
        toks_api = w.run_tokens(rows, max_new_tokens=args.new_tokens, honor_eot=False)
    dd.barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        toks_api = w.run_tokens(rows, max_new_tokens=args.new_tokens, honor_eot=False)
    dd.barrier()
    e2e_step_s = dd.allmax((time.perf_counter() - t1) / args.steps)
    w.close()
    assert toks_api == toks_resident, "the C-API path and the resident path disagree on the token ids"

    # ---- the library's own multi-GPU path: ONE handle on rank 0 that owns an engine per GPU (B200W_DEVICES), host threads,
    # no collective; all N x B chunks through AX_WHISPER_RunPCMTokens; tokens must equal what the ranks produced ---------
    phase("e2e leg done")
    all_toks = dd.gather_tokens(toks_resident, args.new_tokens)
    if not args.no_extras and args.config == 2 and world > 1:
        lib = None
        dd.barrier()
        dd.host_barrier()
        if rank == 0:
            try:
                os.environ["B200W_DEVICES"] = ",".join(str(i) for i in range(world))
                wl = pkg.Whisper(args.arch, util.model_root(args.arch), "zh")
                big = [pcm[i] for i in range(B)]
                keep = []
                for rr in range(1, world):
                    _, a = synth_batch(B, rr * B, pinned=False)
                    keep.append(a)
                    big += [a[i] for i in range(B)]
                wl.run_tokens(big, max_new_tokens=args.new_tokens, honor_eot=False)  # warm: per-GPU workspaces and graphs
                ts = []
                for _ in range(2):
                    t2 = time.perf_counter()
                    got = wl.run_tokens(big, max_new_tokens=args.new_tokens, honor_eot=False)
                    ts.append(time.perf_counter() - t2)
                wl.close()
                ref = [list(map(int, row)) for a in all_toks for row in a]
                bad = [(i, next((k for k in range(min(len(got[i]), len(ref[i]))) if got[i][k] != ref[i][k]), -1)) for i in range(len(ref)) if got[i] != ref[i]]
                lib = {"handle": "one AX_WHISPER handle, B200W_DEVICES=%s, one engine + host thread per GPU" % os.environ["B200W_DEVICES"],
                       "chunks": len(big), "wall_s": min(ts), "audio_s_per_s": CHUNK_S * len(big) / min(ts), "host_buffers": "pageable",
                       "tokens_equal_per_rank_result": not bad}
                if bad:
                    lib["mismatches"] = {"count": len(bad), "per_rank": [sum(1 for i, _ in bad if i // B == rr) for rr in range(world)],
                                         "first": bad[:8], "local_half_equal": got[:B] == toks_resident}
            except Exception as ex:  # the headline line must survive a failing extra
                lib = {"error": "%s: %s" % (type(ex).__name__, ex)}
            finally:
                os.environ.pop("B200W_DEVICES", None)
        dd.host_barrier()  # the other ranks wait on the host: their GPUs belong to rank 0's handle meanwhile
        if rank == 0:
            extra["library_dp"] = lib

    phase("library multi-GPU leg done")
    # ---- the other BASELINE configurations, compact (configs 1, 3 and the long-form 4; config 0 is the CPU report) ------
    if not args.no_extras and args.config == 2:
        others = {}
        for idx in (1, 3, 4):
            a, b, n = CONFIGS[idx]
            if idx == 4:
                lo, hi = shard_range(b, rank, world)
                b = max(hi - lo, 1)
                b = int(dd.allmax(b))  # same batch on every rank (the last ranks of an uneven split pad by one window)
            try:
                line = compact_config_line(pkg, util, dd, a, b, n, pk, first_index=100000 * idx)
                if idx == 4:
                    audio_s = 120 * CHUNK_S
                    line.update({"workload": "1 h = 120 x 30 s windows over %d GPU(s), decoded until the 448-token context is full" % world,
                                 "rtf": (line["ms_per_step"] / 1e3) / audio_s, "audio_s_per_s": audio_s / (line["ms_per_step"] / 1e3)})
                if idx == 1 and world > 1:
                    line["note"] = "configs[1] is a 1-GPU configuration: every rank runs its own 64 chunks here"
                others["config%d" % idx] = line
            except Exception as ex:
                others["config%d" % idx] = {"error": "%s: %s" % (type(ex).__name__, ex)}
        extra["other_configs"] = others

    phase("other configurations done")
    if rank == 0:
        audio_s = CHUNK_S * B * world
        value = audio_s / (step_ms / 1e3)

        def xattn(nb, ms):
            nbytes = nb * 1500 * d * 2 * 2  # K and V of every (sequence, head), bf16
            ach = nbytes / (ms / 1e3) / 1e9
            return {"sequences": nb, "bytes_per_launch": nbytes, "ms_per_launch": ms, "achieved": ach, "frac": ach / pk["hbm_gbs"]}

        in_step, full = xattn(nb_step, x_step_ms), xattn(B, x_full_ms)
        xattn_kernel = "cross_attention_stream_kernel" if nb_step * H >= 296 else "cross_attention_split_kernel"
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(xattn_kernel, {}).get("%s_b%d" % (args.arch, nb_step))
        line = {
            "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {**workload_config(args, world), "l2": "inputs larger than L2 (cross K/V cache %.1f GB per GPU vs 126 MB L2)"
                       % (2 * L * B * 1500 * d * 2 / 1e9)},
            "clocks": clocks,
            "e2e": {"value": audio_s / e2e_step_s, "unit": "audio-s/s", "h2d_bytes_per_step": B * CHUNK_SAMPLES * 4 + B * 4,
                    "d2h_bytes_per_step": B * 448 * 4, "api": "AX_WHISPER_RunPCMTokens (reference-shaped C ABI, pinned host PCM)",
                    "timing": "wall clock around the C-ABI call, barrier + synchronize both sides"},
            "gpu_launches": r["launches"],
            "roofline": {"kernel": xattn_kernel, "bound": "hbm", "achieved": in_step["achieved"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": in_step["frac"], "traffic": traffic,
                         "peak_source": pk["source"] + " (burst copy figure; kernel timed alone; a read-only stream can exceed the read+write copy rate)",
                         "bytes_per_launch": in_step["bytes_per_launch"], "ms_per_launch": in_step["ms_per_launch"],
                         "launch_shape": "%d sequences = one decoder micro-batch, the launch the timed step issues" % nb_step,
                         "whole_batch_launch": full},
            "stages": {**stage_fracs(r, dims, args.arch, B, args.new_tokens, pk), "wall_ms_per_step": wall_step_s * 1e3},
        }
        if not args.no_extras and args.config == 2:
            if world == 1:
                try:
                    extra.setdefault("other_configs", {})["config0"] = config0_report()
                except Exception as ex:
                    extra.setdefault("other_configs", {})["config0"] = {"error": "%s: %s" % (type(ex).__name__, ex)}
            else:
                # a CPU report: it does not depend on N, and the other ranks' processes wait on this one meanwhile (their
                # spinning waits and a 32-thread torch pool on the same cores slow each other down by orders of magnitude)
                extra.setdefault("other_configs", {})["config0"] = {"note": "CPU report of configs[0]: measured in the N = 1 run (python bench.py, or --config 0)"}
        if extra:
            line["extra"] = extra
        if not args.no_cpu_baseline and world == 1:
            import torch as _t

            cores = os.cpu_count() or 1
            _t.set_num_threads(cores)
            n_chunks = 6  # ~15 s of CPU work on this box's host cores
            sec = cpu_reference_pass(args.arch, n_chunks, args.new_tokens, 7000)
            line["cpu_baseline"] = {
                "value": CHUNK_S * n_chunks / sec, "unit": "audio-s/s", "cores": cores,
                "kind": "reference" if util.mel_ref_lib() is not None else "port",
                "sample": "%d x 30 s chunks, Whisper-%s, 4 SOT + %d greedy steps, one pass (%.1f s): mel = %s; encoder/decoder = fp32 torch "
                          "restatement of the reference's exported graphs (onnxruntime not installed)"
                          % (n_chunks, args.arch, args.new_tokens, sec,
                             "the reference's own C++ frontend (oracle/_ref)" if util.mel_ref_lib() is not None else "numpy port")}
        phase("CPU legs done")
        print(json.dumps(line), flush=True)
    if world > 1:
        dd.host_barrier()  # the other ranks wait on the host while rank 0 finishes its CPU legs
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
