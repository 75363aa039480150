"""CPU: the plain-C oracle against the SAME OpenCV routines the reference calls, live through cv2 (skipped without cv2;
tests/test_oracle_golden.py holds the frozen copy).  This is the pin of oracle/ (see oracle/md_oracle.c header)."""
import numpy as np
import pytest

import cvref
from motion_detection_b200 import synth

cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("size", [(53, 37), (640, 480), (121, 77)])
def test_pyramid_scharr_exact(oracle, size):
    w, h = size
    img = np.random.default_rng(w).integers(0, 256, (h, w), dtype=np.uint8)
    lv, pyr = cv2.buildOpticalFlowPyramid(img, (40, 40), 5, None, True)
    mine = oracle.pyramid(img)
    assert len(mine) == lv + 1
    for l in range(lv + 1):
        assert np.array_equal(mine[l], pyr[2 * l])
        assert np.array_equal(oracle.scharr(mine[l]), pyr[2 * l + 1])


def test_lk_matches_cv2(oracle):
    fr, _ = synth.sequence(640, 480, 2, seed=1234)
    pts = oracle.grid_points(640, 480, 10)
    assert np.array_equal(pts, cvref.grid(640, 480, 10))
    p2, st = oracle.lk(fr[0], fr[1], pts)
    q2, qs = cvref.lk(fr[0], fr[1], pts)
    assert (st != qs).mean() < 0.002
    ok = (st == 1) & (qs == 1)
    d = np.linalg.norm(p2[ok] - q2[ok], axis=1)
    assert d.mean() < 1e-3 and (d == 0).mean() > 0.5


@pytest.mark.parametrize("size", [(640, 480), (1920, 1080), (333, 211)])
def test_warp_mask_exact(oracle, size):
    w, h = size
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (h, w), dtype=np.uint8)
    b = rng.integers(0, 256, (h, w), dtype=np.uint8)
    for H in (synth.camera_matrix(w, h, 1, h31=1e-6, h32=-2e-6), np.eye(3),
              np.array([[1, 0, 1 / 64.0], [0, 1, 0.5], [0, 0, 1.0]]),
              np.array([[1.1, 0.2, -30.3], [-0.1, 0.9, 20.7], [1e-4, -2e-4, 1.0]])):
        assert np.array_equal(oracle.warp_perspective(a, H), cv2.warpPerspective(a, H, (w, h)))
        assert np.array_equal(oracle.motion_mask(a, b, H, thresh=60), cvref.mask_chain(a, b, H, thresh=60))


def test_intended_egomotion_agrees_with_findhomography(oracle):
    fr, Hs = synth.sequence(640, 480, 2, seed=1234)
    r = oracle.process_pair(fr[0], fr[1], min_vector_size=0.2)
    inl = r["inlier_mask"] > 0
    assert r["inliers"] == inl.sum() > 2500
    Hc, _ = cv2.findHomography(r["pts"][inl], r["next"][inl], 0)
    assert np.linalg.norm(r["H"] - Hc) / np.linalg.norm(Hc) < 1e-5
    assert np.linalg.norm(r["H"] - Hs[0]) / np.linalg.norm(Hs[0]) < 0.02


def test_varflow_matches_cv2_restatement(oracle):
    fr, _ = synth.sequence(96, 80, 2, seed=5, camera=False, blobs=0, whole_field=(0.75, -0.5), margin=16)
    U, V = oracle.varflow(fr[0], fr[1])
    Uc, Vc = cvref.varflow(fr[0], fr[1])
    assert np.abs(U - Uc).max() < 2e-5 and np.abs(V - Vc).max() < 2e-5
    # the multigrid corrections are numerically inert
    U2, V2 = oracle.varflow(fr[0], fr[1], literal=False)
    assert np.abs(U - U2).max() < 1e-3


def test_fit_subspace_matches_numpy_svd(oracle):
    traj, is_out = synth.trajectories(1500, 5, num_motions=2, seed=3, noise=0.05)
    n, res, cols, outl, thr = oracle.fit_subspace(traj, num_motions=2, sigma=0.5, seed=7)
    T, F, _ = traj.shape
    d = 8
    data = np.zeros((2 * F, T), np.float32)
    data[0::2] = traj[:, :, 0].T
    data[1::2] = traj[:, :, 1].T
    xm = np.float32(data[0].astype(np.float64).sum() / T)
    ym = np.float32(data[1].astype(np.float64).sum() / T)
    data[0::2] = data[0::2] - xm
    data[1::2] = ym - data[1::2]
    D = data.astype(np.float64)
    r = oracle.glibc_rand(7, 50 * d)
    best, bres, bcols = 0, None, None
    for it in range(50):
        c = [r[it * d + k] % T for k in range(d)]
        Uu, S, _ = np.linalg.svd(D[:, c], full_matrices=True)
        if S[-1] < 1e-9 * S[0]:
            continue
        P = np.eye(2 * F) - Uu[:, :d] @ Uu[:, :d].T
        rr = np.abs(np.einsum("it,ij,jt->t", D, P, D))
        cnt = int((rr < (2 * F - d) * 0.25).sum())
        if cnt > best:
            best, bres, bcols = cnt, rr, c
    assert n == best and list(cols) == bcols
    assert np.abs(res - bres).max() < 1e-4
    assert abs(thr - 0.25 * 0.115) < 1e-12      # chi-square p99 table entry 2 (outlier_detector.cpp:22)


def test_line_aa_exact(oracle):
    """oracle.line_aa == cv2.line(..., 1, LINE_AA) on random segments, also partly / wholly outside the image, 1 and 3 channels."""
    rng = np.random.default_rng(11)
    for t in range(1500):
        shp = (40, 50, 3) if t % 2 else (40, 50)
        img = rng.integers(0, 256, shp, dtype=np.uint8)
        ref, got = img.copy(), img.copy()
        p1 = (int(rng.integers(-30, 80)), int(rng.integers(-30, 70)))
        p2 = (p1[0] + int(rng.integers(-45, 46)), p1[1] + int(rng.integers(-45, 46)))
        col = tuple(int(c) for c in rng.integers(0, 256, 3))
        cv2.line(ref, p1, p2, col, 1, cv2.LINE_AA, 0)
        oracle.line_aa(got, p1, p2, col)
        assert np.array_equal(ref, got), (p1, p2, col)


@pytest.mark.parametrize("channels", [1, 3])
def test_show_optical_flow_vectors_exact(oracle, channels):
    rng = np.random.default_rng(5 + channels)
    w, h, ps = 200, 150, 10
    img = rng.integers(0, 256, (h, w, 3) if channels == 3 else (h, w), dtype=np.uint8)
    gy, gx = np.mgrid[0:h:ps, 0:w:ps]
    vec = np.stack([gx.ravel(), gy.ravel(), rng.normal(4, 14, gx.size), rng.normal(-2, 14, gx.size)], axis=1).astype(np.float64)
    vec[::5, 2:] = 0.0                     # vectors below min_vector_size
    vec[3::11, :2] = -1.0; vec[3::11, 2:] = 0.0      # failed points
    col = (10, 200, 90) if channels == 3 else (200,)
    ref, n = cvref.show_optical_flow_vectors(img, vec, ps, 0.2, col)
    got, m = oracle.draw_flow(img, vec, ps, 0.2, col)
    assert n == m and n > 100
    assert np.array_equal(ref, got)
