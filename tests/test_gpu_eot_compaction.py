"""EOT handling of the batched greedy loop (Whisper.cpp:219: stop at EOT): finished sequences are dropped from the decoder's
slot list at the 16-step polls, so the rest of the decode costs what the still-running sequences cost.

Random-init weights practically never emit the real EOT id, so the test model is the `tiny` directory with its config's
"eot" re-pointed to a token id the model DOES emit at different steps in different sequences.  Because a sequence's
arithmetic does not depend on the batch it runs in, the EOT-honouring result must equal, exactly, the EOT-ignoring result
cut at the first occurrence of that id -- with sequences leaving the batch at different times in between."""
import collections
import json
import os
import shutil

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def test_finished_sequences_leave_the_batch(pkg, tmp_path):
    arch, B, n_new = "tiny", 48, 120
    src = os.path.join(util.model_root(arch), arch)
    audios = [util.synth_audio("NUS"[i % 3], 480000 if i % 2 else 200000 + 4000 * i, 7000 + i) for i in range(B)]
    eng = pkg.Engine(util.model_root(arch), arch, 0, B)
    free, _ = eng.transcribe(audios, max_new_tokens=n_new, honor_eot=False)
    eng.close()
    # a token that ends different sequences at different steps (and leaves some running to the end)
    first = collections.defaultdict(dict)
    for b, t in enumerate(free):
        for i, tok in enumerate(t):
            first[tok].setdefault(b, i)
    def spread(tok):
        steps = sorted(first[tok].values())
        return (len(set(s // 16 for s in steps)), len(steps))
    cands = [t for t in first if 4 <= len(first[t]) <= B - 2 and t < 50257]
    assert cands, "the test model emits no token suitable as a stand-in EOT"
    eot = max(cands, key=spread)
    expect = [t[: t.index(eot)] if eot in t else t for t in free]
    lens = sorted(len(t) for t in expect)
    print("stand-in EOT %d: %d of %d sequences stop early, lengths %s" % (eot, sum(len(t) < n_new for t in expect), B, lens))
    # model directory with the re-pointed EOT
    root = str(tmp_path / "models")
    dst = os.path.join(root, arch)
    os.makedirs(dst)
    for f in os.listdir(src):
        if f.endswith("_config.json"):
            cfg = json.load(open(os.path.join(src, f)))
            cfg["eot"] = int(eot)
            json.dump(cfg, open(os.path.join(dst, f), "w"))
        elif not f.startswith("."):
            os.symlink(os.path.join(src, f), os.path.join(dst, f))
    eng = pkg.Engine(root, arch, 0, B)
    assert eng.dims.eot == eot
    c0, _ = eng.decode_stats()
    got, times = eng.transcribe(audios, max_new_tokens=n_new, honor_eot=True)
    c1, active = eng.decode_stats()
    assert got == expect, [(b, len(g), len(e)) for b, (g, e) in enumerate(zip(got, expect)) if g != e]
    print("compactions: %d, sequences still decoding at the end: %d, decoder steps %d" % (c1 - c0, active, times["decode_steps"]))
    assert c1 - c0 >= 1, "finished sequences were never dropped from the batch"
    assert sum(len(t) == n_new for t in expect) <= active < B  # the unfinished ones are all still there, finished ones are gone
    # a second EOT-honouring run (slot map rebuilt from the identity) is deterministic
    got2, _ = eng.transcribe(audios, max_new_tokens=n_new, honor_eot=True)
    assert got2 == got
    # the stateful single-step API still sees slot i == sequence i after a compacted decode
    toks, _ = eng.greedy(B, max_new_tokens=4, honor_eot=False)
    assert [t[:4] for t in free] == toks
    eng.close()
