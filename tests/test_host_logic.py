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
