#!/usr/bin/env python
"""Every entry point of the C ABI once, at small sizes, for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
  compute-sanitizer --tool memcheck python tools/sanitize_run.py
Prints one line per call; exits non-zero when a call fails or a result is not finite."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motion_detection_b200 import capi, synth  # noqa: E402


def main():
    w, h = 320, 240
    frames, _ = synth.sequence(w, h, 6, seed=3, blobs=2)
    ctx = capi.Context(width=w, height=h, max_batch=5, pixel_step=10, min_vector_size=0.2, seed=1, diff_threshold=40)
    r = ctx.process_batch(frames[:4])                            # K1 .. K4, pipelined chunks, TMA paths
    print("process_batch", r["inliers"].tolist(), int((r["mask"] > 0).sum()))
    r = ctx.process_batch(frames[4:], chain=True)
    print("process_batch chained", r["inliers"].tolist())
    rgb = np.repeat(frames[:2, ..., None], 3, axis=3)
    r = ctx.process_batch(rgb)                                   # 8UC3 input (gray conversion fused into level 0)
    print("process_batch rgb", r["inliers"].tolist())
    H = np.array([[1.01, 0.004, -1.3], [-0.003, 0.99, 0.6], [2e-6, -1e-6, 1.0]])
    m = ctx.motion_mask(frames[0], frames[1], H, thresh=40)      # K4 fast path (projective) ...
    H2 = np.array([[0.9, -0.4, 60.0], [0.4, 0.9, -50.0], [0, 0, 1.0]])
    m2 = ctx.motion_mask(frames[0], frames[1], H2, thresh=40)    # ... and the gather path (strong rotation)
    print("motion_mask", int((m > 0).sum()), int((m2 > 0).sum()))
    pts = np.array([[33.3, 40.1], [160.5, 120.25], [300.0, 200.0]], np.float32)
    ctx.process_batch(frames[:2])
    nxt, st = ctx.lk_flow(0, 1, pts)                             # k_lk_tma (arbitrary points)
    print("lk_flow", nxt.round(2).tolist(), st.tolist())
    t = ctx.track_trajectories(frames[:5])
    print("track_trajectories", int((t["len"] == 5).sum()))
    ctx.window_reset()
    for f in frames[:5]:
        ctx.window_push(f)
    d = ctx.window_detect(num_motions=2)
    print("window_detect", {k: (v if np.isscalar(v) else np.asarray(v).shape) for k, v in d.items()})
    traj = t["traj"][t["len"] == 5]
    s = ctx.fit_subspace(traj, num_motions=2)
    print("fit_subspace", s["inliers"])
    rng = np.random.default_rng(0)
    c = ctx.cluster_points(rng.uniform(0, 300, (200, 2)).astype(np.float32))
    print("cluster_points", {k: np.asarray(v).shape for k, v in c.items()} if isinstance(c, dict) else type(c))
    vec = np.concatenate([rng.uniform(0, 300, (100, 2)), rng.normal(0, 3, (100, 2))], axis=1)
    v = ctx.cluster_vectors(vec)
    print("cluster_vectors", type(v).__name__)
    o = ctx.find_outliers(rng.normal(0, 2, (500, 2)))
    print("find_outliers", type(o).__name__)
    ctx.close()
    ctx = capi.Context(width=160, height=120, max_batch=2, pixel_step=10, min_vector_size=0.2, seed=1, flow_engine=capi.MD_FLOW_VARFLOW)
    small, _ = synth.sequence(160, 120, 3, seed=5, blobs=1)
    U, V = ctx.varflow(small[0], small[1])                       # VarFlow (cluster path)
    assert np.isfinite(U).all() and np.isfinite(V).all()
    r = ctx.process_batch(small)                                 # two VarFlow lanes side by side
    print("varflow", float(np.abs(U).max()), r["inliers"].tolist())
    ctx.close()
    ctx = capi.Context(width=160, height=120, max_batch=1, vf_grid_barrier=1)
    U2, V2 = ctx.varflow(small[0], small[1])                     # cooperative grid-barrier path
    assert np.array_equal(U, U2) and np.array_equal(V, V2)
    ctx.close()
    print("sanitize_run ok")


if __name__ == "__main__":
    main()
