"""Oracle parity ON THE BASELINE ARCHITECTURES AND BATCH SIZES, on the code path bench.py times.

For Whisper-small B=256 (configs[2]), Whisper-base B=64 (configs[1]) and Whisper-turbo B=128 (configs[3]) the whole batch
goes through the product kernels -- CTA-pair tcgen05 GEMMs over 128-chunk encoder sub-batches, the tcgen05 encoder
attention, two decoder micro-batches on two streams, the streaming cross-attention kernel, 8-step CUDA graphs -- and two full
30 s chunks, one from each micro-batch / encoder sub-batch, are compared with the fp32 CPU oracle:

  (i)   cross K/V element-wise (reference encoder graph, export_onnx.py:193-213): max-abs error <= 2 % of max|ref|,
        cosine >= 0.999 per tensor;
  (ii)  teacher-forced logits (decoder graph, export_onnx.py:312-387) for 24 generated tokens: EVERY step is compared,
        max-abs error <= 1.5 % of max|logit|; per step the arg-max equals the oracle's unless the oracle's top-2 margin is
        below twice the measured error;
  (iii) the CUDA-graph path (what bench.py runs) produces, under the same teacher forcing, exactly the arg-max sequence of
        the eager path whose logits were compared -- for all B sequences;
  (iv)  decoder alone on the ORACLE's cross K/V (loaded through b200w_set_cross_kv, the a8 contract): logits within the
        same tolerance with the encoder's bf16 error taken out.
The measured numbers are written to gpurun_out/parity_r02_<arch>.json (committed copy: profiles/parity_r02.json)."""
import json
import os

import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu

N_NEW = 24
CASES = [("base", 64), ("small", 256), ("turbo", 128)]


def _dump(arch, rec):
    out = os.path.join(util.ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_r02_%s.json" % arch), "w") as f:
        json.dump(rec, f, indent=1)


@pytest.mark.parametrize("arch,B", CASES)
def test_bench_path_parity(pkg, arch, B):
    eng = pkg.Engine(util.model_root(arch), arch, 0, B)
    oracle = util.load_oracle(arch)
    base = [util.synth_audio("NUS"[i % 3], 480000 if i % 4 != 2 else 300000 + 5000 * i, 900 + 10 * i + len(arch)) for i in range(8)]
    audios = [base[i % 8] for i in range(B)]
    sel = [3, B // 2 + 5]  # one full 30 s chunk in each micro-batch (and, for B = 256, in each 128-chunk encoder sub-batch)
    assert all(len(audios[i]) == 480000 for i in sel)
    rec = dict(arch=arch, batch=B, selected=sel, n_new=N_NEW)

    mel = eng.logmel(audios)
    ref_mel = util.reference_mel([audios[i] for i in sel], oracle.n_mels)
    rec["mel_max_abs_err"] = float(np.abs(mel[sel] - ref_mel).max())
    assert rec["mel_max_abs_err"] <= util.MEL_TOL
    eng.encoder(batch=B, return_cross=False)

    # (i) cross K/V of the selected chunks, element-wise
    with torch.no_grad():
        rk, rv = oracle.encoder(ref_mel)  # [L, 2, 1500, d]
    rk, rv = rk.numpy(), rv.numpy()
    rec["cross_kv"] = []
    for j, i in enumerate(sel):
        ck, cv = eng.get_cross_kv(i, 1)
        for name, got, ref in (("cross_k", ck[:, 0], rk[:, j]), ("cross_v", cv[:, 0], rv[:, j])):
            rel = float(np.abs(got - ref).max() / np.abs(ref).max())
            cos = util.cosine(got, ref)
            print("%s B=%d chunk %d %s: max-abs-err/max|ref| %.4f cosine %.6f" % (arch, B, i, name, rel, cos))
            rec["cross_kv"].append(dict(chunk=i, tensor=name, rel_err=rel, cosine=cos))
            assert rel <= util.ENC_REL_TOL and cos >= util.ENC_COS_TOL, (name, i, rel, cos)

    # (ii) teacher-forced logits, every step, for the selected chunks inside the full batch
    ref = oracle.greedy(torch.from_numpy(rk), torch.from_numpy(rv), max_new_tokens=N_NEW, honor_eot=False, keep_logits=True)
    ref_logits = np.stack(ref["logits"])  # [N_NEW, 2, V]
    margins = np.stack(ref["top2_margin"])
    forced = np.zeros((B, N_NEW), np.int32)
    for i in range(B):
        forced[i] = ref["tokens"][0 if i < B // 2 else 1]
    toks_eager, logits = eng.greedy(B, max_new_tokens=N_NEW, honor_eot=False, forced_tokens=forced, keep_logits=True, logit_rows=sel)
    err_steps = np.abs(logits[:N_NEW] - ref_logits).max(axis=2)  # [N_NEW, 2]
    tol = util.logit_tol(ref_logits)
    rec["logits"] = dict(max_abs_err=float(err_steps.max()), max_abs_ref=float(np.abs(ref_logits).max()), tol=tol,
                         per_step_max_abs_err=[float(x) for x in err_steps.max(axis=1)], steps_compared=int(err_steps.size))
    print("%s B=%d teacher-forced logits: %d (step, sequence) pairs compared, max-abs err %.4f, max|logit| %.2f, tol %.4f"
          % (arch, B, err_steps.size, err_steps.max(), np.abs(ref_logits).max(), tol))
    assert err_steps.max() <= tol
    flips = []
    for s in range(N_NEW):
        for j, i in enumerate(sel):
            if toks_eager[i][s] != ref["tokens"][j][s]:
                flips.append(dict(step=s, chunk=i, margin=float(margins[s][j])))
                assert margins[s][j] < 2 * err_steps.max() + 1e-6, "arg-max differs at a confident step: margin %.4f" % margins[s][j]
    rec["argmax"] = dict(compared=N_NEW * len(sel), identical=N_NEW * len(sel) - len(flips), low_margin_flips=flips,
                         min_margin=float(margins.min()))
    print("%s B=%d arg-max: %d of %d steps identical to the oracle (low-margin flips: %s)" % (arch, B, N_NEW * 2 - len(flips), N_NEW * 2, flips))

    # (iii) the CUDA-graph path = the path bench.py times: same arg-max sequence as the eager path, all B sequences
    toks_graph, _ = eng.greedy(B, max_new_tokens=N_NEW, honor_eot=False, forced_tokens=forced)
    assert toks_graph == toks_eager
    # copies of a chunk in other slots of the batch (both micro-batches) give the same arg-max sequence
    for i in range(B):
        if i % 8 == sel[0] % 8 and i < B // 2:
            assert toks_graph[i] == toks_graph[sel[0]]
        if i % 8 == sel[1] % 8 and i >= B // 2:
            assert toks_graph[i] == toks_graph[sel[1]]
    rec["graph_equals_eager_sequences"] = B

    # free-running (no forcing) on the graph path, against the oracle's free-running tokens with the margin rule
    toks_free, _ = eng.greedy(B, max_new_tokens=N_NEW, honor_eot=False)
    rep = util.token_report([toks_free[i] for i in sel], ref["tokens"], ref["top2_margin"], 2 * tol)
    rec["free_running"] = rep
    print("%s B=%d free-running: %s" % (arch, B, rep))
    assert all(r["ok"] for r in rep)
    eng.close()

    # (iv) decoder alone on the oracle's cross K/V (a8: caches supplied by the caller)
    eng2 = pkg.Engine(util.model_root(arch), arch, 0, 2)
    eng2.set_cross_kv(rk, rv)
    gk, gv = eng2.get_cross_kv(0, 2)
    assert np.abs(gk - rk).max() <= 2 ** -8 * np.abs(rk).max() and np.abs(gv - rv).max() <= 2 ** -8 * np.abs(rv).max()  # bf16 rounding only
    _, lg = eng2.greedy(2, max_new_tokens=N_NEW, honor_eot=False, forced_tokens=np.array(ref["tokens"], np.int32), keep_logits=True)
    err2 = float(np.abs(lg[:N_NEW] - ref_logits).max())
    rec["decoder_on_oracle_cross_kv"] = dict(max_abs_err=err2, tol=tol)
    print("%s decoder on the oracle's cross K/V: logits max-abs err %.4f (tol %.4f)" % (arch, err2, tol))
    assert err2 <= tol
    eng2.close()
    _dump(arch, rec)
