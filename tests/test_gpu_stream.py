"""GPU: long chained streams (ROS-free emulation of the node's frame loop): the pyramid ring wraps many
times, every pair is still the oracle's pair, and the per-stream counters add up."""
import numpy as np
import pytest

from motion_detection_b200 import synth

pytestmark = pytest.mark.gpu


def test_long_chained_stream_ring_wraparound(capi, oracle):
    w, h, n = 320, 240, 121
    frames, _ = synth.sequence(w, h, n, seed=77, blobs=2)
    ctx = capi.Context(width=w, height=h, max_batch=7, pixel_step=10, min_vector_size=0.2, seed=100)
    masks, Hs, nv = [], [], []
    res = ctx.process_batch(frames[:6])                         # pairs 0..4
    masks.append(res["mask"]); Hs.append(res["H"]); nv.append(res["num_vectors"])
    f = 6
    rng = np.random.default_rng(0)
    while f < n:
        k = int(min(n - f, rng.integers(1, 8)))                 # ragged batch sizes 1..7 through a ring of 8 slots
        res = ctx.process_batch(frames[f:f + k], chain=True)
        masks.append(res["mask"]); Hs.append(res["H"]); nv.append(res["num_vectors"])
        f += k
    masks = np.concatenate(masks); Hs = np.concatenate(Hs); nv = np.concatenate(nv)
    assert len(masks) == n - 1
    for p in range(0, n - 1, 17):                               # sampled pairs against the oracle (seed = seed0 + pair index)
        ref = oracle.process_pair(frames[p], frames[p + 1], min_vector_size=0.2, seed=100 + p)
        assert abs(int(nv[p]) - ref["num_vectors"]) <= 2
        assert np.linalg.norm(Hs[p] - ref["H"]) / np.linalg.norm(ref["H"]) < 1e-4, p
        assert (masks[p] == ref["mask"]).mean() >= 0.999, p
    st = ctx.stats()
    assert st["pairs"] == n - 1
    assert st["mask_pixels"] == int((masks > 0).sum())


def test_skip_frames_like_the_node(capi, oracle):
    # imageCallback keeps every `skip_frames`-th image (node.cpp:237-261): the hot path only ever sees the kept frames
    frames, _ = synth.sequence(320, 240, 9, seed=5, blobs=1)
    kept = frames[::2]
    ctx = capi.Context(width=320, height=240, max_batch=4, pixel_step=10, min_vector_size=0.4, seed=1)
    res = ctx.process_batch(kept)
    for p in range(len(kept) - 1):
        ref = oracle.process_pair(kept[p], kept[p + 1], min_vector_size=0.4, seed=1 + p)
        assert np.linalg.norm(res["H"][p] - ref["H"]) / np.linalg.norm(ref["H"]) < 1e-4
        assert (res["mask"][p] == ref["mask"]).mean() >= 0.999


def test_two_contexts_interleaved_on_one_gpu(capi):
    """Several camera streams per GPU (streams.shard_streams with more streams than GPUs): contexts share nothing -- calls
    interleaved between two contexts give exactly what each context gives alone."""
    w, h = 320, 240
    fa, _ = synth.sequence(w, h, 9, seed=11, blobs=2)
    fb, _ = synth.sequence(w, h, 9, seed=12, blobs=1)

    def alone(frames, seed):
        ctx = capi.Context(width=w, height=h, max_batch=4, pixel_step=10, min_vector_size=0.2, seed=seed)
        r1 = ctx.process_batch(frames[:5])
        r2 = ctx.process_batch(frames[5:], chain=True)
        ctx.close()
        return r1, r2

    a1, a2 = alone(fa, 3)
    b1, b2 = alone(fb, 4)
    ca = capi.Context(width=w, height=h, max_batch=4, pixel_step=10, min_vector_size=0.2, seed=3)
    cb = capi.Context(width=w, height=h, max_batch=4, pixel_step=10, min_vector_size=0.2, seed=4)
    x1 = ca.process_batch(fa[:5])
    y1 = cb.process_batch(fb[:5])
    x2 = ca.process_batch(fa[5:], chain=True)
    y2 = cb.process_batch(fb[5:], chain=True)
    for got, ref in ((x1, a1), (x2, a2), (y1, b1), (y2, b2)):
        for k in ("next", "status", "keep", "H", "num_vectors", "inliers", "mask"):
            assert np.array_equal(got[k], ref[k]), k


@pytest.mark.parametrize("lanes", [2, 3])
def test_batch_pipeline_equals_one_chained_context(capi, lanes):
    """streams.BatchPipeline: batches of one sequence alternate over `lanes` contexts as asynchronous host-buffer calls
    (MD_MEM_HOST_ASYNC, unchained with a one-frame overlap, md_set_pair_index before each).  Every output of every pair -- the RANSAC
    draws included -- must equal a single context's chained calls bit for bit."""
    from motion_detection_b200 import streams
    w, h, B, nb = 320, 240, 4, 7
    frames, _ = synth.sequence(w, h, B * nb + 1, seed=31, blobs=2)
    kw = dict(width=w, height=h, max_batch=B, pixel_step=10, min_vector_size=0.2, seed=9)
    ctx = capi.Context(**kw)
    ref = [ctx.process_batch(frames[:B + 1])]
    for k in range(1, nb):
        ref.append(ctx.process_batch(frames[1 + k * B:1 + (k + 1) * B], chain=True))
    pipe = streams.BatchPipeline(lanes=lanes, **kw)
    P = pipe.ctxs[0].P
    bufs = [dict(next=np.zeros((B, P, 2), np.float32), status=np.zeros((B, P), np.uint8), keep=np.zeros((B, P), np.uint8),
                 H=np.zeros((B, 3, 3), np.float64), num_vectors=np.zeros(B, np.int32), inliers=np.zeros(B, np.int32),
                 mask=np.zeros((B, h, w), np.uint8)) for _ in range(lanes)]
    outs = [capi.MdOutputs(capi._ptr(b["next"]), capi._ptr(b["status"]), capi._ptr(b["keep"]), capi._ptr(b["H"]), capi._ptr(b["num_vectors"]),
                           capi._ptr(b["inliers"]), capi._ptr(b["mask"]), w, w * h) for b in bufs]
    got = []

    def collect(ticket):
        pipe.wait(ticket)
        got.append({k: v.copy() for k, v in bufs[pipe.lane_of(ticket)].items()})

    for k in range(nb):
        if k >= lanes:
            collect(k - lanes)                                   # the lane's buffers are free again
        fr = frames[k * B:k * B + B + 1]
        assert fr.flags["C_CONTIGUOUS"]
        assert pipe.submit(fr.ctypes.data, 1, w, w * h, B + 1, outs[k % lanes]) == k
    for k in range(max(0, nb - lanes), nb):
        collect(k)
    assert len(got) == nb
    for k in range(nb):
        for key in ("next", "status", "keep", "H", "num_vectors", "inliers", "mask"):
            assert np.array_equal(got[k][key], ref[k][key]), (k, key)
    assert pipe.stats()["pairs"] == B * nb
    pipe.close()
    ctx.close()
