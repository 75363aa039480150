"""CPU: host-side logic that needs no GPU -- synthetic sequences, stream sharding, the gloo statistics gather."""
import os
import subprocess
import sys

import numpy as np

from motion_detection_b200 import streams, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synthetic_sequence_is_seeded_and_thresholdable():
    a, Ha = synth.sequence(160, 120, 3, seed=9)
    b, _ = synth.sequence(160, 120, 3, seed=9)
    c, _ = synth.sequence(160, 120, 3, seed=10)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert a.dtype == np.uint8 and a.shape == (3, 120, 160) and Ha.shape == (2, 3, 3)
    bg = a[0][a[0] < 255]
    assert bg.max() <= 60 and (a[0] == 255).sum() > 50      # background in [0,60], movers at 255 (threshold is 190)


def test_shard_streams_round_robin():
    assert streams.shard_streams(8, 8, 3) == [3]
    assert streams.shard_streams(8, 2, 1) == [1, 3, 5, 7]
    assert streams.shard_streams(3, 4, 3) == []
    allr = sorted(sum((streams.shard_streams(10, 4, r) for r in range(4)), []))
    assert allr == list(range(10))


def test_shard_sequence_pairs_overlap_one_frame():
    # pair mode: every pair (k, k+1) is owned by exactly one rank; chunks overlap by one frame
    chunks = [streams.shard_sequence(101, 4, r) for r in range(4)]
    pairs = []
    for lo, hi in chunks:
        pairs += [(k, k + 1) for k in range(lo, hi - 1)]
    assert pairs == [(k, k + 1) for k in range(100)]
    assert all(chunks[i][1] - 1 == chunks[i + 1][0] for i in range(3))
    # trajectory mode: F-1 frames of overlap
    chunks = [streams.shard_sequence(101, 4, r, overlap=4) for r in range(4)]
    assert all(chunks[i][1] - 4 == chunks[i + 1][0] for i in range(3))


def test_stats_gather_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, json\n"
        "sys.path.insert(0, %r)\n"
        "import torch.distributed as dist\n"
        "from motion_detection_b200 import streams\n"
        "dist.init_process_group('gloo')\n"
        "r = dist.get_rank()\n"
        "mine = dict(pairs=10 + r, mask_pixels=100 * (r + 1), tracked=7, inliers=5 + r)\n"
        "out = streams.gather_stats(mine)\n"
        "assert len(out) == 2 and out[0]['pairs'] == 10 and out[1]['pairs'] == 11 and out[1]['mask_pixels'] == 200\n"
        "assert streams.shard_streams(4, dist.get_world_size(), r) == [r, r + 2]\n"
        "t = streams.max_over_ranks(1.5 + r)\n"
        "assert abs(t - 2.5) < 1e-9\n"
        "dist.destroy_process_group()\n" % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


def test_batch_pipeline_bookkeeping(monkeypatch):
    """streams.BatchPipeline without a GPU: batches alternate over the lanes, every lane's pair counter is set to the sequence's pair
    index before its batch, a lane is waited for before it is fed again, every call is an unchained MD_MEM_HOST_ASYNC batch."""
    from motion_detection_b200 import capi
    log = []

    class FakeCtx:
        n = 0

        def __init__(self, device=0, **kw):
            self.id = FakeCtx.n
            FakeCtx.n += 1
            self.kw = kw

        def set_pair_index(self, i):
            log.append(("idx", self.id, i))

        def raw_process_batch(self, ptr, ch, pitch, stride, count, chain, outputs, mem):
            assert not chain and mem == capi.MD_MEM_HOST_ASYNC
            log.append(("run", self.id, ptr, count))

        def sync(self):
            log.append(("sync", self.id))

        def stats(self):
            return dict(pairs=3, mask_pixels=1, tracked=2, inliers=1, kernel_launches=4, lk_iterations=5, lk_levels=6, graph_replays=7,
                        last_H=None, device=0)

        def close(self):
            log.append(("close", self.id))

    monkeypatch.setattr(capi, "Context", FakeCtx)
    pipe = streams.BatchPipeline(lanes=2, width=64, height=48)
    assert [c.kw for c in pipe.ctxs] == [dict(width=64, height=48)] * 2
    t = [pipe.submit(1000 + k, 1, 64, 64 * 48, 5, None) for k in range(3)]      # 4 pairs per batch
    assert t == [0, 1, 2] and pipe.pairs == 12 and [pipe.lane_of(x) for x in t] == [0, 1, 0]
    assert log == [("idx", 0, 0), ("run", 0, 1000, 5), ("idx", 1, 4), ("run", 1, 1001, 5),
                   ("sync", 0), ("idx", 0, 8), ("run", 0, 1002, 5)]              # lane 0 is waited for before batch 2
    pipe.wait(0)                                                                # already retired: nothing happens
    assert log[-1][0] == "run"
    pipe.wait(1)
    assert log[-1] == ("sync", 1)
    pipe.drain()
    assert log[-1] == ("sync", 0)
    assert pipe.stats()["pairs"] == 6 and pipe.stats()["graph_replays"] == 14
    pipe.close()
    assert log[-2:] == [("close", 0), ("close", 1)]
