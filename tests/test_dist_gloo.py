"""world_size-2 gloo test (CPU) of the N>1 host logic: contiguous sharding, the single all-gather
of per-rank partial results and the combine step, with the CPU oracle standing in for the
per-rank engine (the GPU engine itself is covered by tests/test_gpu_msm.py::test_msm_sharded_partial_combine)."""
import os
import random
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, q):
    import ctypes as C
    import numpy as np
    import oracle_lib
    import pyref
    from curve25519_dalek_b200.sharding import shard_range, all_gather_windows
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle_lib.load()
    rnd = random.Random(42)                       # same inputs on every rank
    B = orc.basepoint()
    ts = [rnd.randrange(pyref.L) for _ in range(n)]
    ss = [rnd.randrange(pyref.L) for _ in range(n)]
    lo, hi = shard_range(n, rank, world)
    pts = [orc.scalarmul(ts[i].to_bytes(32, "little"), B) for i in range(lo, hi)]
    part = orc.msm("optional", [ss[i].to_bytes(32, "little") for i in range(lo, hi)], pts) if hi > lo else orc.identity()
    limbs = np.array(orc.p3_limbs(part), dtype=np.uint64)      # stand-in for the window accumulators
    gathered = all_gather_windows(limbs.tobytes(), world, dist=dist)
    allp = np.frombuffer(gathered, dtype=np.uint64).reshape(world, 20)
    total = orc.identity()
    for r in range(world):
        total = orc.add(total, orc.p3_from_limbs([int(x) for x in allp[r]]))
    k = sum(a * b for a, b in zip(ss, ts)) % pyref.L
    want = orc.compress(orc.scalarmul(k.to_bytes(32, "little"), B))
    q.put((rank, orc.compress(total) == want, (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    from curve25519_dalek_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 1 << 20, (1 << 24) + 3):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_sharded_msm_over_gloo():
    world, n = 2, 41
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    assert all(ok for _, ok, _ in res)
    assert sorted(r for _, _, r in res) == [(0, 21), (21, 41)]
