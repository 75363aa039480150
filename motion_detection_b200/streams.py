"""Multi-GPU plumbing: the path shards over independent units (camera streams, or chunks of one long sequence), one
process per GPU, with NO collective on the per-frame path.  torch.distributed (NCCL on GPUs, gloo in CPU tests) is
used only for the barrier, the max-over-ranks timing and the gather of per-stream statistics at report time
(SURVEY.md section 8e)."""


def shard_streams(n_streams, world_size, rank):
    """Round-robin assignment of independent camera streams to ranks (one md_ctx per stream)."""
    return list(range(rank, n_streams, world_size))


def shard_sequence(n_frames, world_size, rank, overlap=1):
    """Contiguous chunk [lo, hi) of one long sequence for `rank`.  Consecutive chunks share `overlap` frames:
    1 in pair mode (a pair needs only its two frames, optical_flow_calculator.cpp:30), F-1 in trajectory mode (:161)."""
    units = n_frames - overlap            # pairs (or windows) to distribute
    base, rem = divmod(units, world_size)
    lo = rank * base + min(rank, rem)
    cnt = base + (1 if rank < rem else 0)
    return lo, lo + cnt + overlap


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def gather_stats(mine):
    """All-gather a small dict of integer per-stream counters; returns the list over ranks."""
    import torch
    dist = _dist()
    keys = sorted(mine)
    if dist is None:
        return [dict(mine)]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([int(mine[k]) for k in keys], dtype=torch.int64, device=dev)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [dict(zip(keys, o.tolist())) for o in out]


def max_over_ranks(value):
    import torch
    dist = _dist()
    if dist is None:
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
