#!/usr/bin/env python
"""ms per md_draw_flow call (showOpticalFlowVectors on the device), 1080p, device-resident image and batch outputs, beside the
oracle's sequential cv::line loop on one host core.  usage: python tools/bench_draw.py"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from motion_detection_b200 import capi, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    w, h, ps = 1920, 1080, 10
    frames, _ = synth.sequence(w, h, 2, seed=1234)
    ctx = capi.Context(width=w, height=h, max_batch=1, pixel_step=ps, min_vector_size=0.2, seed=1)
    r = ctx.process_batch(frames)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    rgb = np.repeat(frames[1][..., None], 3, axis=2)
    d_img = torch.from_numpy(rgb).to(dev)
    d_out = torch.empty_like(d_img)
    d_next = torch.from_numpy(r["next"][0]).to(dev); d_st = torch.from_numpy(r["status"][0]).to(dev); d_keep = torch.from_numpy(r["keep"][0]).to(dev)
    col = torch.tensor([255, 0, 0], dtype=torch.uint8)
    d_n = torch.zeros(1, dtype=torch.int32, device=dev)

    def call():
        rc = capi.lib().md_draw_flow(ctx._h, C.c_void_p(d_img.data_ptr()), 3, 3 * w, None, 0, C.c_void_p(d_next.data_ptr()),
                                     C.c_void_p(d_st.data_ptr()), C.c_void_p(d_keep.data_ptr()), C.c_void_p(col.data_ptr()),
                                     C.c_void_p(d_out.data_ptr()), 3 * w, C.c_void_p(d_n.data_ptr()), capi.MD_MEM_DEVICE)
        assert rc == 0
    for _ in range(5):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(50):
        call()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    vec = O.flow_field_row_major(O.grid_points(w, h, ps), r["next"][0], r["status"][0], r["keep"][0])
    t0 = time.perf_counter()
    ref, n = O.draw_flow(rgb, vec, ps, 0.2, (255, 0, 0))
    cpu_ms = 1e3 * (time.perf_counter() - t0)
    same = int((d_out.cpu().numpy() != ref).any(axis=2).sum())
    print(json.dumps({"what": "md_draw_flow 1920x1080x3, %d arrows, device-resident" % int(d_n.item()), "ms_per_call": ms,
                      "gbs_algorithmic": 2 * 3 * w * h / (ms * 1e-3) / 1e9, "cpu_oracle_ms_1_thread": cpu_ms, "pixels_differing_from_oracle": same}))


if __name__ == "__main__":
    main()
