"""Host-side sharding of an MSM that spans several GPUs: contiguous shards of the pair
range, ONE exchange of the per-rank window accumulators, combine on every rank.

`ShardedMsm` is the device-resident form (used by bench.py at N > 1): the shard's MSM, the all-gather of
the records and the combine are all enqueued on the engine's stream; the host blocks once, for the 192-byte
result.  `all_gather_windows` is the plain host form of the same exchange (CPU tensors: the gloo test)."""


def shard_range(n_total, rank, world):
    """Contiguous [lo, hi) of the pair range owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_size(n_total, world):
    """Size of the largest shard: the `n_shard` every rank passes to the engine (it selects the window width)."""
    return (n_total + world - 1) // world


def all_gather_windows(windows_bytes, world, dist=None, device=None):
    """All-gather each rank's window accumulators (bytes: nwin x 160) over torch.distributed.
    Returns the rank-major concatenation as bytes (world x nwin x 160).  Point addition is not a
    reduction operator NCCL knows, so the exchange is gather-then-add, not all-reduce."""
    import torch
    if world == 1:
        return bytes(windows_bytes)
    if dist is None:
        import torch.distributed as dist
    mine = torch.frombuffer(bytearray(windows_bytes), dtype=torch.uint8)
    if device is not None:
        mine = mine.to(device)
    out = torch.empty(world * mine.numel(), dtype=torch.uint8, device=mine.device)
    dist.all_gather_into_tensor(out, mine)
    return out.cpu().numpy().tobytes()


class ShardedMsm:
    """One rank's side of an MSM over `world` GPUs (one process per GPU, torch.distributed initialised with NCCL
    when world > 1).  All device work of a call is ordered on the engine's own stream, which torch sees as an
    ExternalStream: NCCL's all-gather waits for the partial MSM and the combine waits for the all-gather through
    CUDA events, never through the host."""

    def __init__(self, engine, world, n_shard, device):
        import torch
        self.torch, self.eng, self.world, self.n_shard = torch, engine, world, n_shard
        self.rec_bytes = engine.msm_partial_bytes(n_shard)
        self.mine = torch.zeros(self.rec_bytes, dtype=torch.uint8, device=device)
        self.gathered = torch.zeros(world * self.rec_bytes, dtype=torch.uint8, device=device) if world > 1 else self.mine
        self.stream = torch.cuda.ExternalStream(engine.stream_ptr(), device=device)
        torch.cuda.synchronize(device)

    def run(self, scalars, points, n_local, point_fmt, device_ptrs):
        """Returns (rc, compressed result); rc 1 == None (some shard held an undecodable point)."""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            self.eng.edwards_msm_partial_async(scalars, points, n_local, self.n_shard, self.mine.data_ptr(),
                                               point_fmt=point_fmt, device_ptrs=device_ptrs)
            if self.world > 1:
                import torch.distributed as dist
                dist.all_gather_into_tensor(self.gathered, self.mine)     # the one exchange step
            rc, comp, _ = self.eng.edwards_msm_combine_dev(self.gathered.data_ptr(), self.world, self.n_shard)
        return rc, comp
