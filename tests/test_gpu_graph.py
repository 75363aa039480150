"""GPU: the captured CUDA graph of a batch (md_config.cuda_graphs, the default) replays the eager launch sequence bit for bit --
device-resident buffers, chained batches with pinned host buffers, seeds that advance with the pair counter, the ring position
that is brought back to slot 0 before every chained call -- and pageable host buffers quietly stay on the eager path."""
import numpy as np
import pytest

from motion_detection_b200 import synth

pytestmark = pytest.mark.gpu

KEYS = ("next", "status", "keep", "H", "nv", "inl", "mask")


def _device_outputs(capi, torch, B, P, h, w, dev):
    t = dict(next=torch.zeros((B, P, 2), dtype=torch.float32, device=dev), status=torch.zeros((B, P), dtype=torch.uint8, device=dev),
             keep=torch.zeros((B, P), dtype=torch.uint8, device=dev), H=torch.zeros((B, 9), dtype=torch.float64, device=dev),
             nv=torch.zeros((B,), dtype=torch.int32, device=dev), inl=torch.zeros((B,), dtype=torch.int32, device=dev),
             mask=torch.zeros((B, h, w), dtype=torch.uint8, device=dev))
    outs = capi.MdOutputs(t["next"].data_ptr(), t["status"].data_ptr(), t["keep"].data_ptr(), t["H"].data_ptr(), t["nv"].data_ptr(),
                          t["inl"].data_ptr(), t["mask"].data_ptr(), w, w * h)
    return t, outs


def _snapshot(torch, t):
    torch.cuda.synchronize()
    return {k: t[k].cpu().numpy().copy() for k in KEYS}


@pytest.mark.parametrize("size,B", [((320, 240), 6), ((640, 480), 4)])
def test_graph_replay_equals_eager_device_buffers(capi, size, B):
    import torch
    w, h = size
    dev = torch.device("cuda", 0)
    clips = [synth.sequence(w, h, B + 1, seed=40 + i, blobs=2)[0] for i in range(2)]
    results = {}
    for graphs in (0, 1):
        ctx = capi.Context(width=w, height=h, max_batch=B, pixel_step=10, min_vector_size=0.2, seed=9, cuda_graphs=graphs)
        P = ctx.P
        d_frames = [torch.from_numpy(c).to(dev) for c in clips]
        t, outs = _device_outputs(capi, torch, B, P, h, w, dev)
        got = []
        # the same two resident clips over and over: each buffer is seen once (eager), captured on its second call, replayed after
        for call in range(8):
            f = d_frames[call % 2]
            ctx.raw_process_batch(f.data_ptr(), 1, w, w * h, B + 1, False, outs, capi.MD_MEM_DEVICE)
            got.append(_snapshot(torch, t))
        st = ctx.stats()
        results[graphs] = (got, st)
        ctx.close()
    eager, graph = results[0], results[1]
    assert eager[1]["graph_replays"] == 0
    assert graph[1]["graph_replays"] >= 4, graph[1]
    assert graph[1]["kernel_launches"] > 0
    for call in range(8):
        for k in KEYS:
            assert np.array_equal(eager[0][call][k], graph[0][call][k]), (call, k)
    # the seeds advance with the pair counter: the same clip gives another RANSAC draw two calls later (same flow, same vectors)
    assert np.array_equal(graph[0][0]["next"], graph[0][2]["next"])
    assert graph[1]["pairs"] == 8 * B


def test_graph_replay_chained_pinned_host_buffers(capi, oracle):
    import torch
    w, h, B = 320, 240, 5
    frames, _ = synth.sequence(w, h, B + 1, seed=3, blobs=2)
    fwd = torch.from_numpy(frames[1:].copy()).pin_memory()
    bwd = torch.from_numpy(frames[:B][::-1].copy()).pin_memory()
    first = torch.from_numpy(frames.copy()).pin_memory()
    results = {}
    for graphs in (0, 1):
        ctx = capi.Context(width=w, height=h, max_batch=B, pixel_step=10, min_vector_size=0.2, seed=21, cuda_graphs=graphs)
        P = ctx.P
        pin = dict(next=torch.zeros((B, P, 2), dtype=torch.float32).pin_memory(), status=torch.zeros((B, P), dtype=torch.uint8).pin_memory(),
                   keep=torch.zeros((B, P), dtype=torch.uint8).pin_memory(), H=torch.zeros((B, 9), dtype=torch.float64).pin_memory(),
                   nv=torch.zeros((B,), dtype=torch.int32).pin_memory(), inl=torch.zeros((B,), dtype=torch.int32).pin_memory(),
                   mask=torch.zeros((B, h, w), dtype=torch.uint8).pin_memory())
        outs = capi.MdOutputs(pin["next"].data_ptr(), pin["status"].data_ptr(), pin["keep"].data_ptr(), pin["H"].data_ptr(),
                              pin["nv"].data_ptr(), pin["inl"].data_ptr(), pin["mask"].data_ptr(), w, w * h)
        got = []
        ctx.raw_process_batch(first.data_ptr(), 1, w, w * h, B + 1, False, outs, capi.MD_MEM_HOST)       # f0 .. fB
        got.append({k: pin[k].numpy().copy() for k in KEYS})
        for call in range(8):                                                                              # back and forth through the clip
            # backward: (fB, fB-1) ... (f1, f0); forward: (f0, f1) ... (fB-1, fB)
            src = bwd if call % 2 == 0 else fwd
            ctx.raw_process_batch(src.data_ptr(), 1, w, w * h, B, True, outs, capi.MD_MEM_HOST)
            got.append({k: pin[k].numpy().copy() for k in KEYS})
        results[graphs] = (got, ctx.stats())
        ctx.close()
    eager, graph = results[0], results[1]
    assert eager[1]["graph_replays"] == 0
    assert graph[1]["graph_replays"] >= 4, graph[1]
    for call in range(9):
        for k in KEYS:
            assert np.array_equal(eager[0][call][k], graph[0][call][k]), (call, k)
    # and the replayed pairs are still the oracle's pairs: forward pass number 2 (call index 4 in `got`), pair 0 = (f0, f1)
    ref = oracle.process_pair(frames[0], frames[1], min_vector_size=0.2, seed=21 + B + 3 * B)
    r = graph[0][4]
    assert np.linalg.norm(r["H"][0].reshape(3, 3) - ref["H"]) / np.linalg.norm(ref["H"]) < 1e-4
    assert (r["mask"][0] == ref["mask"]).mean() >= 0.999


def test_pageable_host_buffers_stay_eager(capi):
    w, h, B = 320, 240, 3
    frames, _ = synth.sequence(w, h, B + 1, seed=8, blobs=1)
    ctx = capi.Context(width=w, height=h, max_batch=B, pixel_step=10, min_vector_size=0.2, seed=2)
    a = ctx.process_batch(frames)
    b = ctx.process_batch(frames)
    c = ctx.process_batch(frames)
    assert np.array_equal(a["next"], b["next"]) and np.array_equal(b["next"], c["next"])
    assert np.array_equal(a["mask"].shape, c["mask"].shape)
    assert ctx.stats()["pairs"] == 3 * B
    ctx.close()
