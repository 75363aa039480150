"""Seeded synthetic sequences for parity tests and the bench.

numpy only.  Background = blurred uniform noise scaled to [0, 60]; movers are discs / patches at 255 so
that |warp(prev) - cur| can exceed the reference's fixed threshold of 190
(common/src/optical_flow_calculator.cpp:127).
"""
import math

import numpy as np


def _blur(img, sigma):
    r = int(math.ceil(4 * sigma))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    k /= k.sum()
    pad = np.pad(img, ((0, 0), (r, r)), mode="reflect")
    out = np.zeros_like(img)
    for i, kv in enumerate(k):
        out += kv * pad[:, i:i + img.shape[1]]
    pad = np.pad(out, ((r, r), (0, 0)), mode="reflect")
    out2 = np.zeros_like(img)
    for i, kv in enumerate(k):
        out2 += kv * pad[i:i + img.shape[0], :]
    return out2


def texture(w, h, seed, margin=128, sigma=2.0, lo=0.0, hi=60.0):
    """Blurred-noise canvas of size (h + 2*margin, w + 2*margin), min-max normalised to [lo, hi] (float64)."""
    rng = np.random.default_rng(seed)
    t = rng.random((h + 2 * margin, w + 2 * margin))
    t = _blur(t, sigma)
    t = (t - t.min()) / (t.max() - t.min())
    return lo + (hi - lo) * t


def _sample_bilinear(canvas, xs, ys):
    x0 = np.floor(xs).astype(np.int64)
    y0 = np.floor(ys).astype(np.int64)
    fx = xs - x0
    fy = ys - y0
    H, W = canvas.shape
    x0 = np.clip(x0, 0, W - 2)
    y0 = np.clip(y0, 0, H - 2)
    a = canvas[y0, x0]
    b = canvas[y0, x0 + 1]
    c = canvas[y0 + 1, x0]
    d = canvas[y0 + 1, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def camera_matrix(w, h, k, rot_deg=0.05, scale=1.0005, tx=1.2, ty=-0.8, h31=0.0, h32=0.0):
    """Cumulative camera map after k frames: maps frame-0 pixel coords to frame-k pixel coords
    (rotation + scale about the image centre, translation, optional projective terms)."""
    cx, cy = w / 2.0, h / 2.0
    a = math.radians(rot_deg)
    A = np.array([[scale * math.cos(a), -scale * math.sin(a), 0.0],
                  [scale * math.sin(a), scale * math.cos(a), 0.0],
                  [h31, h32, 1.0]])
    T0 = np.array([[1, 0, -cx], [0, 1, -cy], [0, 0, 1.0]])
    T1 = np.array([[1, 0, cx + tx], [0, 1, cy + ty], [0, 0, 1.0]])
    step = T1 @ A @ T0
    M = np.eye(3)
    for _ in range(k):
        M = step @ M
    return M / M[2, 2]


def sequence(w, h, n_frames, seed=1234, camera=True, blobs=3, h31=0.0, h32=0.0, patch=False, whole_field=None,
             margin=128):
    """Returns (frames [n][h][w] u8, H_true [n-1][3][3] f64 mapping frame k -> frame k+1).

    camera=True  : C2/C3 style global affine (+ optional projective) camera motion, `blobs` discs at 255.
    camera=False : C1 style static camera; patch=True adds a textured 1/4-size patch in [200, 255]
                   translating by (0.75, -0.5) px/frame; whole_field=(vx, vy) translates everything (C1b).
    """
    canvas = texture(w, h, seed, margin=margin)
    rng = np.random.default_rng(seed + 7919)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    discs = []
    for i in range(blobs):
        r = [20, 30, 40][i % 3] * max(w, h) / 1920.0
        r = max(r, 6.0)
        cx = rng.uniform(0.2 * w, 0.8 * w)
        cy = rng.uniform(0.2 * h, 0.8 * h)
        sp = rng.uniform(2.0, 4.0)
        ang = rng.uniform(0, 2 * math.pi)
        discs.append((cx, cy, r, sp * math.cos(ang), sp * math.sin(ang)))
    ptex = None
    if patch:
        ptex = texture(w // 4, h // 4, seed + 1, margin=0, lo=200.0, hi=255.0)
    frames = np.empty((n_frames, h, w), np.uint8)
    Hs = []
    for k in range(n_frames):
        if camera:
            M = camera_matrix(w, h, k, h31=h31, h32=h32)
        elif whole_field is not None:
            M = np.array([[1, 0, whole_field[0] * k], [0, 1, whole_field[1] * k], [0, 0, 1.0]])
        else:
            M = np.eye(3)
        Mi = np.linalg.inv(M)
        den = Mi[2, 0] * xx + Mi[2, 1] * yy + Mi[2, 2]
        sx = (Mi[0, 0] * xx + Mi[0, 1] * yy + Mi[0, 2]) / den + margin
        sy = (Mi[1, 0] * xx + Mi[1, 1] * yy + Mi[1, 2]) / den + margin
        img = _sample_bilinear(canvas, sx, sy)
        if ptex is not None:
            px0 = 0.3 * w + 0.75 * k
            py0 = 0.4 * h - 0.5 * k
            u = xx - px0
            v = yy - py0
            inside = (u >= 0) & (v >= 0) & (u < ptex.shape[1] - 1) & (v < ptex.shape[0] - 1)
            pv = _sample_bilinear(ptex, np.clip(u, 0, ptex.shape[1] - 1.001), np.clip(v, 0, ptex.shape[0] - 1.001))
            img = np.where(inside, pv, img)
        for (cx, cy, r, vx, vy) in discs:
            m = (xx - (cx + vx * k)) ** 2 + (yy - (cy + vy * k)) ** 2 <= r * r
            img = np.where(m, 255.0, img)
        frames[k] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
        if k > 0:
            if camera:
                Hk = camera_matrix(w, h, 1, h31=h31, h32=h32)
            elif whole_field is not None:
                Hk = np.array([[1, 0, whole_field[0]], [0, 1, whole_field[1]], [0, 0, 1.0]])
            else:
                Hk = np.eye(3)
            Hs.append(Hk)
    return frames, np.array(Hs)


def trajectories(T, F, num_motions=2, seed=1, noise=0.05, outlier_frac=0.1, w=640, h=480):
    """Synthetic point trajectories [T][F][2] f32: `num_motions` independent affine motions of 3-D-ish point
    sets (rank-4 each) plus `outlier_frac` random-walk outliers. Returns (traj, is_outlier)."""
    rng = np.random.default_rng(seed)
    traj = np.zeros((T, F, 2), np.float64)
    is_out = np.zeros(T, bool)
    labels = rng.integers(0, num_motions, T)
    P3 = np.concatenate([rng.uniform(0, w, (T, 1)), rng.uniform(0, h, (T, 1)), rng.uniform(-50, 50, (T, 1)),
                         np.ones((T, 1))], axis=1)
    for m in range(num_motions):
        sel = labels == m
        for f in range(F):
            ang = 0.01 * f * (m + 1)
            A = np.array([[math.cos(ang), -math.sin(ang), 0.02 * f * (m + 1), 1.5 * f * (m + 1)],
                          [math.sin(ang), math.cos(ang), -0.015 * f, -1.0 * f * (m + 1)]])
            traj[sel, f, :] = P3[sel] @ A.T
    n_out = int(outlier_frac * T)
    out_idx = rng.choice(T, n_out, replace=False)
    is_out[out_idx] = True
    for f in range(1, F):
        traj[out_idx, f, :] = traj[out_idx, f - 1, :] + rng.normal(0, 6.0, (n_out, 2))
    traj += rng.normal(0, noise, traj.shape)
    return traj.astype(np.float32), is_out
