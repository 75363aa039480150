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


class BatchPipeline:
    """One camera sequence over `lanes` contexts of ONE GPU, fed alternately with MD_MEM_HOST_ASYNC batches (include/motion_b200.h).

    A host-buffer md_process_batch call has a head (first upload, first pyramids, planes and window sums) and a tail (last egomotion
    fit, last mask, last download, the host's wait) during which the flow kernels -- 70 % of the work -- have nothing to run.  With
    two contexts the head and tail of batch k run beside the flow kernels of batch k +- 1.  Every batch is handed in UNCHAINED: `count`
    frames, the first being the last frame of the previous batch (a pair needs only its two frames, optical_flow_calculator.cpp:30;
    one extra upload and pyramid per batch), and the lane's pair counter is set to the sequence's pair index first, so the results --
    RANSAC draws included -- are those of a single chained context (tests/test_gpu_stream.py).

    submit() returns a ticket; wait(ticket) returns when that batch's outputs are in the caller's (pinned) buffers.  A lane's buffers
    (frames and outputs of its batch in flight) must stay untouched until its ticket has been waited for; submit() itself waits for
    the lane's previous batch."""

    def __init__(self, lanes=2, device=0, **ctx_kw):
        from . import capi
        self._capi = capi
        self.ctxs = [capi.Context(device=device, **ctx_kw) for _ in range(lanes)]
        self.batches = 0                       # batches submitted
        self.pairs = 0                         # pairs submitted = index of the next pair of the sequence
        self._pending = [None] * lanes         # ticket in flight per lane

    def lane_of(self, ticket):
        return ticket % len(self.ctxs)

    def submit(self, frames_ptr, channels, pitch, frame_stride, count, outputs):
        """count frames (count - 1 pairs) at frames_ptr, `outputs` an MdOutputs of host pointers; returns the ticket."""
        ticket = self.batches
        lane = self.lane_of(ticket)
        if self._pending[lane] is not None:
            self.wait(self._pending[lane])
        ctx = self.ctxs[lane]
        ctx.set_pair_index(self.pairs)
        ctx.raw_process_batch(frames_ptr, channels, pitch, frame_stride, count, False, outputs, self._capi.MD_MEM_HOST_ASYNC)
        self._pending[lane] = ticket
        self.batches += 1
        self.pairs += count - 1
        return ticket

    def wait(self, ticket):
        lane = self.lane_of(ticket)
        if self._pending[lane] == ticket:
            self.ctxs[lane].sync()
            self._pending[lane] = None

    def drain(self):
        for t in sorted(p for p in self._pending if p is not None):
            self.wait(t)

    def stats(self):
        """Counters summed over the lanes."""
        out = {}
        for c in self.ctxs:
            st = c.stats()
            for k in ("pairs", "mask_pixels", "tracked", "inliers", "kernel_launches", "lk_iterations", "lk_levels", "graph_replays"):
                out[k] = out.get(k, 0) + int(st[k])
        return out

    def close(self):
        self.drain()
        for c in self.ctxs:
            c.close()
