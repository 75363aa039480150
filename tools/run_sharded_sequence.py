#!/usr/bin/env python
"""ONE long sequence split into contiguous chunks with a 1-frame overlap, one chunk per GPU, no collective on
the frame path -- and the proof that the split is invisible: every pair's egomotion, vector count and mask checksum equal the
ones a single context computes over the whole sequence (pair i draws its RANSAC samples with seed0 + i on either path).

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/run_sharded_sequence.py --frames 65
"""
import argparse
import json
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_chunk(capi, frames, seed0, first_pair, device, batch):
    """pairs first_pair .. first_pair + len(frames) - 2 of the sequence, in chained batches; returns per-pair records"""
    h, w = frames.shape[1:]
    ctx = capi.Context(width=w, height=h, max_batch=batch, pixel_step=10, min_vector_size=0.2, seed=seed0 + first_pair, device=device)
    out = []
    f, first = 0, True
    while f < len(frames) - (0 if not first else 1):
        n = min(batch + (1 if first else 0), len(frames) - f)
        res = ctx.process_batch(frames[f:f + n], chain=not first)
        for b in range(len(res["H"])):
            out.append((res["H"][b].tobytes(), int(res["num_vectors"][b]), int(res["inliers"][b]), zlib.crc32(res["mask"][b].tobytes())))
        f += n
        first = False
    ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=65)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--batch", type=int, default=8)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from motion_detection_b200 import capi, streams, synth
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    frames, _ = synth.sequence(a.width, a.height, a.frames, seed=1234)
    lo, hi = streams.shard_sequence(a.frames, world, rank, overlap=1)
    t0 = time.perf_counter()
    mine = run_chunk(capi, frames[lo:hi], 100, lo, local, a.batch)
    dt = time.perf_counter() - t0
    # gather (pair index, crc of everything) -- a few bytes per pair, off the frame path
    crc = torch.tensor([[lo + i, zlib.crc32(r[0]) ^ r[1] ^ (r[2] << 8) ^ r[3]] for i, r in enumerate(mine)], dtype=torch.int64, device="cuda")
    counts = torch.tensor([len(mine)], dtype=torch.int64, device="cuda")
    if world > 1:
        all_counts = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(all_counts, counts)
        mx = int(max(c.item() for c in all_counts))
        pad = torch.full((mx, 2), -1, dtype=torch.int64, device="cuda")
        pad[:len(mine)] = crc
        allc = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(allc, pad)
        got = {int(r[0]): int(r[1]) for t in allc for r in t.cpu().numpy() if r[0] >= 0}
    else:
        got = {int(r[0]): int(r[1]) for r in crc.cpu().numpy()}
    if rank == 0:
        ref = run_chunk(capi, frames, 100, 0, local, a.batch)
        want = {i: zlib.crc32(r[0]) ^ r[1] ^ (r[2] << 8) ^ r[3] for i, r in enumerate(ref)}
        same = sum(1 for i in want if got.get(i) == want[i])
        print(json.dumps({"tool": "run_sharded_sequence", "n_gpus": world, "frames": a.frames, "pairs": len(want), "pairs_gathered": len(got),
                          "pairs_identical_to_single_context": same, "chunk_of_rank0": [lo, hi], "seconds_rank0_chunk": dt}), flush=True)
        assert same == len(want) == len(got)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
