"""Host-side sharding helpers for an MSM that spans several GPUs: contiguous shards of
the pair range, one exchange of the per-rank window accumulators, combine on every rank."""


def shard_range(n_total, rank, world):
    """Contiguous [lo, hi) of the pair range owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


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
