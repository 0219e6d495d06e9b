"""The data-parallel host logic of bench.py / multi-GPU use, exercised with world_size = 2 over gloo on CPU: every rank
owns a contiguous shard of the utterance list, there is no data-path collective, only a barrier and a MAX-reduce of the
per-rank time."""
import os
import subprocess
import sys

import util

WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(sys.argv[1], "tests")); sys.path.insert(0, sys.argv[1])
import util, bench
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n_items = 7
lo, hi = bench.shard_range(n_items, rank, world)
# per-rank work on its own shard only (a checksum of each synthetic utterance stands in for the GPU pass)
local = [float(np.abs(util.synth_audio("N", 1600, 2000 + i)).sum()) for i in range(lo, hi)]
t = torch.tensor([0.1 * (rank + 1)], dtype=torch.float64)
dist.barrier()
dist.all_reduce(t, op=dist.ReduceOp.MAX)
gathered = [None] * world
dist.all_gather_object(gathered, (lo, hi, local))
# bench.py's host-side wait (a gloo sub-group: ranks whose GPUs are lent to rank 0's library handle wait without spinning on them)
dd = bench.Dist(rank, 0, world)
assert dd.cpu_group is not None
dd.host_barrier()
# strong split of configs[2]: 256 chunks over 1 / 2 / 4 / 8 ranks, long-form: 120 windows over 8
for w in (1, 2, 4, 8):
    parts = [bench.shard_range(256, r, w) for r in range(w)]
    assert parts[0][0] == 0 and parts[-1][1] == 256 and all(a[1] == b[0] for a, b in zip(parts, parts[1:])) and all(h - l == 256 // w for l, h in parts)
assert [h - l for l, h in (bench.shard_range(120, r, 8) for r in range(8))] == [15] * 8
if rank == 0:
    assert abs(t.item() - 0.1 * world) < 1e-12
    covered = []
    for l, h, vals in gathered:
        assert len(vals) == h - l
        covered += list(range(l, h))
    assert covered == list(range(n_items)), covered
    flat = [v for _, _, vals in gathered for v in vals]
    ref = [float(np.abs(util.synth_audio("N", 1600, 2000 + i)).sum()) for i in range(n_items)]
    assert flat == ref
    print("SHARDING_OK")
dist.destroy_process_group()
"""


def test_two_rank_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", str(script), util.ROOT]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "SHARDING_OK" in out.stdout


def test_shard_range_partitions():
    import bench
    for n in (0, 1, 7, 256, 120):
        for world in (1, 2, 4, 8):
            spans = [bench.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1
