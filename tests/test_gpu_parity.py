"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): integer stages bit-exact given identical inputs; flow <= 0.01 px mean end-point
error; egomotion parameters <= 1e-4 relative; motion masks >= 99.9 % pixel agreement.
"""
import numpy as np
import pytest

from motion_detection_b200 import synth

pytestmark = pytest.mark.gpu

FLOW_EPE_TOL = 0.01          # px, mean end-point error (north_star)
H_REL_TOL = 1e-4             # relative Frobenius error of H (h33 = 1)  (north_star)
MASK_AGREE = 0.999           # fraction of identical mask pixels (north_star)


@pytest.fixture(scope="module")
def seq640():
    return synth.sequence(640, 480, 5, seed=1234)


def _ctx(capi, w, h, **kw):
    kw.setdefault("min_vector_size", 0.2)
    return capi.Context(width=w, height=h, **kw)


# ---- K0 ---------------------------------------------------------------------------------------------------------
def test_gray_bit_exact(capi, oracle):
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    ctx = _ctx(capi, 131, 97)
    assert np.array_equal(ctx.gray(rgb), oracle.gray(rgb))


# ---- K1 ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size", [(640, 480), (333, 211), (1920, 1080), (97, 131)])
def test_pyramid_and_scharr_bit_exact(capi, oracle, size):
    w, h = size
    rng = np.random.default_rng(w * 7 + h)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    ctx = _ctx(capi, w, h, max_batch=2)
    ref = oracle.pyramid(img)
    assert ctx.levels == len(ref)
    for slot in (0, 2):
        ctx.pyramid(img, slot)
        for lvl, r in enumerate(ref):
            assert np.array_equal(ctx.pyramid_read(slot, lvl), r), (slot, lvl)
            assert np.array_equal(ctx.pyramid_read_deriv(slot, lvl), oracle.scharr(r)), (slot, lvl)


# ---- K2 ---------------------------------------------------------------------------------------------------------
def test_lk_flow_matches_oracle_640(capi, oracle, seq640):
    frames, _ = seq640
    ctx = _ctx(capi, 640, 480, max_batch=1)
    ctx.pyramid(frames[0], 0)
    ctx.pyramid(frames[1], 1)
    nxt, st = ctx.lk_flow(0, 1)
    pts = ctx.grid_points()
    assert np.array_equal(pts, oracle.grid_points(640, 480, 10))
    ref, rst = oracle.lk(frames[0], frames[1], pts)
    assert (st != rst).mean() < 0.002
    ok = (st == 1) & (rst == 1)
    d = np.linalg.norm(nxt[ok] - ref[ok], axis=1)
    assert d.mean() < FLOW_EPE_TOL
    assert np.median(d) < 1e-4
    # failed points keep the reference's bookkeeping: same positions where both failed at level 0
    both_fail = (st == 0) & (rst == 0)
    if both_fail.any():
        assert np.linalg.norm(nxt[both_fail] - ref[both_fail], axis=1).mean() < 0.05


def test_lk_flow_explicit_points_and_borders(capi, oracle, seq640):
    frames, _ = seq640
    ctx = _ctx(capi, 640, 480, max_batch=1, pixel_step=4)
    ctx.pyramid(frames[1], 0)
    ctx.pyramid(frames[2], 1)
    rng = np.random.default_rng(5)
    pts = np.concatenate([rng.uniform(-5, 645, (500, 1)), rng.uniform(-5, 485, (500, 1))], axis=1).astype(np.float32)
    pts[:4] = [[0, 0], [639, 479], [0.5, 479.5], [639.25, 0.75]]
    nxt, st = ctx.lk_flow(0, 1, pts)
    ref, rst = oracle.lk(frames[1], frames[2], pts)
    assert (st != rst).mean() < 0.01
    ok = (st == 1) & (rst == 1)
    assert np.linalg.norm(nxt[ok] - ref[ok], axis=1).mean() < FLOW_EPE_TOL


@pytest.mark.parametrize("size,ps", [((640, 480), 10), ((640, 480), 7), ((640, 480), 8), ((333, 211), 3), ((1920, 1080), 10),
                                     ((320, 240), 1), ((640, 480), 4), ((97, 131), 2)])
def test_lk_grid_phase_planes_bit_identical_to_point_kernel(capi, size, ps):
    """Grid mode (pts = NULL) runs k_lk_phase on precomputed sub-pixel phase planes; handing the SAME grid in as explicit
    points runs the per-point window build of k_lk_tma.  Both are the same integer arithmetic: results must be identical
    to the last bit, for every phase-class layout (even / odd / power-of-two pixel_step, dense)."""
    w, h = size
    frames, _ = synth.sequence(w, h, 2, seed=99 + ps)
    ctx = _ctx(capi, w, h, max_batch=1, pixel_step=ps)
    ctx.pyramid(frames[0], 0)
    ctx.pyramid(frames[1], 1)
    pts = ctx.grid_points()
    a, sa = ctx.lk_flow(0, 1)
    b, sb = ctx.lk_flow(0, 1, pts)
    assert np.array_equal(sa, sb)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), np.abs(a - b).max()
    assert sa.mean() > 0.5


def test_lk_flow_vs_cv2_direct(capi, seq640):
    cvref = pytest.importorskip("cvref")
    if not cvref.have_cv2():
        pytest.skip("cv2 missing")
    frames, _ = seq640
    ctx = _ctx(capi, 640, 480, max_batch=1)
    ctx.pyramid(frames[0], 0)
    ctx.pyramid(frames[1], 1)
    nxt, st = ctx.lk_flow(0, 1)
    ref, rst = cvref.lk(frames[0], frames[1], ctx.grid_points())
    ok = (st == 1) & (rst == 1)
    assert (st != rst).mean() < 0.002
    assert np.linalg.norm(nxt[ok] - ref[ok], axis=1).mean() < FLOW_EPE_TOL


# ---- K3 ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [1, 2, 0])
def test_egomotion_identical_inputs(capi, oracle, seq640, mode):
    frames, _ = seq640
    pts = oracle.grid_points(640, 480, 10)
    p2, st = oracle.lk(frames[0], frames[1], pts)
    nv, keep, _ = oracle.flow_filter(pts, p2, st, 0.2)
    ctx = _ctx(capi, 640, 480)
    for seed in (1, 7, 12345):
        ref_n, ref_H, ref_inl = oracle.fit_egomotion(pts, p2, keep, 640, 480, mode=mode, seed=seed)
        got = ctx.fit_egomotion(pts, p2, status=st, mode=mode, seed=seed)
        assert got["num_vectors"] == nv
        assert got["inliers"] == ref_n
        if mode != 0:
            assert np.array_equal(got["inlier_mask"], ref_inl)
        rel = np.linalg.norm(got["H"] - ref_H) / np.linalg.norm(ref_H)
        assert rel < 1e-9, (mode, seed, rel)
        # explicit keep flags give the same answer
        got2 = ctx.fit_egomotion(pts, p2, keep=keep, mode=mode, seed=seed)
        assert np.array_equal(got2["H"], got["H"])


def test_egomotion_first4_nondegenerate(capi, oracle):
    # strict mode (optical_flow_calculator.cpp:120) on four non-collinear leading vectors
    w, h = 160, 120
    ctx = _ctx(capi, w, h, pixel_step=50)
    src = np.array([[0, 0], [0, 50], [50, 100], [100, 0], [50, 50], [100, 100]], np.float32)
    Ht = np.array([[1.01, 0.02, 1.5], [-0.01, 0.99, -0.7], [1e-5, -2e-5, 1.0]])
    q = (Ht @ np.c_[src, np.ones(len(src))].T).T
    dst = (q[:, :2] / q[:, 2:]).astype(np.float32)
    got = ctx.fit_egomotion(src, dst, keep=np.ones(len(src), np.uint8), mode=0)
    n, Href, _ = oracle.fit_egomotion(src, dst, np.ones(len(src), np.uint8), w, h, mode=0)
    assert n == 4 and got["inliers"] == 4
    assert np.linalg.norm(got["H"] - Href) / np.linalg.norm(Href) < 1e-12
    assert np.linalg.norm(got["H"] - Ht) / np.linalg.norm(Ht) < 1e-4


def test_egomotion_too_few_vectors(capi):
    ctx = _ctx(capi, 160, 120, pixel_step=50)
    src = np.array([[0, 0], [0, 50], [50, 0]], np.float32)
    got = ctx.fit_egomotion(src, src + 1, keep=np.ones(3, np.uint8), mode=1)
    assert got["inliers"] == 0 and got["num_vectors"] == 3
    assert np.array_equal(got["H"], np.eye(3))


# ---- K4 ---------------------------------------------------------------------------------------------------------
H_CASES = [
    np.eye(3),
    np.array([[1, 0, 1 / 64.0], [0, 1, 0.5], [0, 0, 1.0]]),
    np.array([[1.0005, -0.00087, 1.25], [0.00087, 1.0005, -1.2], [0, 0, 1.0]]),
    np.array([[1.1, 0.2, -30.3], [-0.1, 0.9, 20.7], [1e-4, -2e-4, 1.0]]),
    np.array([[0.5, 0, 400.0], [0, 0.5, -300.0], [0, 0, 1.0]]),
]


@pytest.mark.parametrize("size", [(640, 480), (333, 211), (1920, 1080)])
def test_fused_mask_bit_exact(capi, oracle, size):
    w, h = size
    frames, _ = synth.sequence(w, h, 2, seed=99 + w, blobs=3)
    rng = np.random.default_rng(w)
    noisy = rng.integers(0, 256, (h, w), dtype=np.uint8)
    ctx = _ctx(capi, w, h)
    for H in H_CASES:
        for prev, cur in ((frames[0], frames[1]), (noisy, frames[1])):
            for morph in (True, False):
                got = ctx.motion_mask(prev, cur, H, morph=morph)
                ref = oracle.motion_mask(prev, cur, H, morph=morph)
                assert np.array_equal(got, ref), (size, H.tolist(), morph)
    # lower threshold so that plenty of pixels fire near the borders as well
    got = ctx.motion_mask(noisy, frames[1], H_CASES[2], thresh=40)
    assert np.array_equal(got, oracle.motion_mask(noisy, frames[1], H_CASES[2], thresh=40))
    assert (got > 0).sum() > 1000


def test_fused_mask_vs_cv2(capi, seq640):
    cvref = pytest.importorskip("cvref")
    if not cvref.have_cv2():
        pytest.skip("cv2 missing")
    frames, _ = seq640
    ctx = _ctx(capi, 640, 480)
    for H in H_CASES:
        assert np.array_equal(ctx.motion_mask(frames[0], frames[1], H), cvref.mask_chain(frames[0], frames[1], H))


def test_mask_properties_full_size(capi):
    # size-independent properties at 1080p: identical frames + identity -> empty; inverted frame -> opening of a constant
    w, h = 1920, 1080
    frames, _ = synth.sequence(w, h, 1, seed=3, blobs=3)
    ctx = _ctx(capi, w, h)
    assert ctx.motion_mask(frames[0], frames[0], np.eye(3)).sum() == 0
    black = np.zeros_like(frames[0])
    white = np.full_like(frames[0], 255)
    assert (ctx.motion_mask(black, white, np.eye(3)) == 255).all()
    # shifting by an integer translation: the mask of (prev, shift(prev)) under that translation is empty inside
    sh = np.zeros_like(frames[0]); sh[:, 7:] = frames[0][:, :-7]
    m = ctx.motion_mask(frames[0], sh, np.array([[1, 0, 7.0], [0, 1, 0], [0, 0, 1]]))
    assert m[:, 7:].sum() == 0


# ---- chain ------------------------------------------------------------------------------------------------------
def _check_pair(res, b, ref, strict_mask=True):
    st, rst = res["status"][b], ref["status"]
    assert (st != rst).mean() < 0.002
    ok = (st == 1) & (rst == 1)
    epe = np.linalg.norm(res["next"][b][ok] - ref["next"][ok], axis=1).mean()
    assert epe < FLOW_EPE_TOL, epe
    assert abs(int(res["num_vectors"][b]) - ref["num_vectors"]) <= 0.002 * len(st) + 1
    relH = np.linalg.norm(res["H"][b] - ref["H"]) / np.linalg.norm(ref["H"])
    assert relH < H_REL_TOL, relH
    agree = (res["mask"][b] == ref["mask"]).mean()
    assert agree >= MASK_AGREE, agree


def test_process_batch_matches_oracle(capi, oracle, seq640):
    frames, Hs = seq640
    ctx = _ctx(capi, 640, 480, max_batch=4, seed=11)
    res = ctx.process_batch(frames)       # 4 pairs in one call
    for b in range(4):
        ref = oracle.process_pair(frames[b], frames[b + 1], min_vector_size=0.2, seed=11 + b)
        _check_pair(res, b, ref)
        # and the egomotion is the true camera motion
        assert np.linalg.norm(res["H"][b] - Hs[b]) / np.linalg.norm(Hs[b]) < 0.02
        assert (res["mask"][b] > 0).sum() > 50


def test_process_batch_chained_equals_unchained(capi, seq640):
    frames, _ = seq640
    a = _ctx(capi, 640, 480, max_batch=4, seed=3)
    full = a.process_batch(frames)
    b = _ctx(capi, 640, 480, max_batch=2, seed=3)
    r0 = b.process_batch(frames[:3])               # pairs 0,1
    r1 = b.process_batch(frames[3:5], chain=True)  # pairs 2,3 from the cached pyramid of frame 2
    for k in ("next", "status", "keep", "H", "num_vectors", "inliers", "mask"):
        got = np.concatenate([r0[k], r1[k]])
        assert np.array_equal(got, full[k]), k
    with pytest.raises(capi.MotionB200Error):
        _ctx(capi, 640, 480, max_batch=2).process_batch(frames[:2], chain=True)
    st = b.stats()
    assert st["pairs"] == 4 and st["mask_pixels"] == int((full["mask"] > 0).sum())
    assert st["tracked"] == int(full["status"].sum()) and st["inliers"] == int(full["inliers"].sum())


def test_process_batch_rgb_input(capi, oracle, seq640):
    frames, _ = seq640
    rgb = np.repeat(frames[:2, :, :, None], 3, axis=3)      # gray replicated: weights sum to 2^15 -> same gray
    ctx = _ctx(capi, 640, 480, max_batch=1, seed=5)
    r3 = ctx.process_batch(rgb)
    ctx2 = _ctx(capi, 640, 480, max_batch=1, seed=5)
    r1 = ctx2.process_batch(frames[:2])
    for k in ("next", "status", "H", "mask"):
        assert np.array_equal(r3[k], r1[k]), k


def test_no_vectors_gives_empty_mask(capi):
    # node default min_vector_size = 1.0 (node.cpp:44) zeroes every sub-pixel vector of a static camera
    frames, _ = synth.sequence(320, 240, 2, seed=8, camera=False, blobs=2)
    ctx = capi.Context(width=320, height=240, max_batch=1, min_vector_size=50.0)
    res = ctx.process_batch(frames)
    assert res["num_vectors"][0] == 0 and res["inliers"][0] == 0
    assert res["mask"].sum() == 0 and np.array_equal(res["H"][0], np.eye(3))


def test_process_batch_1080p(capi, oracle):
    frames, Hs = synth.sequence(1920, 1080, 3, seed=1234)
    ctx = _ctx(capi, 1920, 1080, max_batch=2, seed=21)
    res = ctx.process_batch(frames)
    for b in range(2):
        ref = oracle.process_pair(frames[b], frames[b + 1], min_vector_size=0.2, seed=21 + b)
        _check_pair(res, b, ref)


@pytest.mark.parametrize("size,ps", [((640, 480), 7), ((640, 480), 16), ((320, 240), 3), ((160, 120), 1), ((1280, 720), 12)])
def test_process_batch_other_grid_steps(capi, oracle, size, ps):
    """The chain at other pixel_step values: other phase-class layouts and lattice steps (7, 3: steps that do not divide the
    40-row window; 16: a single class on every level; 1: dense) through planes + window sums + LK + fit + mask."""
    w, h = size
    frames, _ = synth.sequence(w, h, 3, seed=31 + ps, blobs=2)
    ctx = _ctx(capi, w, h, max_batch=2, seed=5, pixel_step=ps)
    res = ctx.process_batch(frames)
    for b in range(2):
        ref = oracle.process_pair(frames[b], frames[b + 1], pixel_step=ps, min_vector_size=0.2, seed=5 + b)
        _check_pair(res, b, ref)


def test_affine_mode_chain(capi, oracle, seq640):
    frames, _ = seq640
    ctx = _ctx(capi, 640, 480, max_batch=1, seed=2, ego_mode=2)
    res = ctx.process_batch(frames[:2])
    ref = oracle.process_pair(frames[0], frames[1], min_vector_size=0.2, seed=2, mode=oracle.MODE_RANSAC_AFFINE)
    _check_pair(res, 0, ref)
    assert abs(res["H"][0][2, 0]) == 0 and abs(res["H"][0][2, 1]) == 0


# ---- trajectories -----------------------------------------------------------------------------------------------
def test_trajectories_match_oracle_bookkeeping(capi, oracle, seq640):
    frames, _ = seq640
    F = 5
    ctx = _ctx(capi, 640, 480, max_batch=F - 1)
    got = ctx.track_trajectories(frames[:F])
    # oracle: calculateOpticalFlowTrajectory bookkeeping (cpp:161-254) around the oracle LK
    pts = oracle.grid_points(640, 480, 10)
    cur = pts.copy()
    traj = [[tuple(p)] for p in pts]
    for j in range(F - 1):
        nxt, st = oracle.lk(frames[j], frames[j + 1], cur)
        new = cur.copy()
        for i in range(len(pts)):
            if st[i] and nxt[i, 0] > 10.0 and nxt[i, 1] > 10.0 and nxt[i, 0] < 640 - 10 and nxt[i, 1] < 480 - 10:
                new[i] = nxt[i]
                traj[i].append(tuple(nxt[i]))
        cur = new
    ref_len = np.array([len(t) for t in traj])
    assert (got["len"] != ref_len).mean() < 0.005
    full = (got["len"] == F) & (ref_len == F)
    assert full.sum() > 0.8 * len(pts)
    ref_full = np.array([traj[i] for i in np.nonzero(full)[0]], np.float32)
    d = np.linalg.norm(got["traj"][full] - ref_full, axis=2)
    assert d.mean() < FLOW_EPE_TOL


# ---- fitSubspace (outlier_detector.cpp:236-331) -----------------------------------------------------------------
@pytest.mark.parametrize("T,F,nm,sigma", [(2000, 5, 2, 0.5), (517, 5, 1, 4.0), (3000, 7, 3, 1.0)])
def test_fit_subspace_matches_oracle(capi, oracle, T, F, nm, sigma):
    traj, _ = synth.trajectories(T, F, num_motions=nm, seed=T, noise=0.05)
    ctx = _ctx(capi, 640, 480)
    for seed in (1, 99):
        n, res, cols, outl, thr = oracle.fit_subspace(traj, num_motions=nm, sigma=sigma, seed=seed)
        got = ctx.fit_subspace(traj, num_motions=nm, sigma=sigma, seed=seed)
        assert got["inliers"] == n
        assert np.array_equal(got["best_cols"], cols)
        assert np.array_equal(got["outlier"], outl)
        assert np.allclose(got["residual"], res, rtol=1e-6, atol=1e-9)
    # explicit column lists instead of rand()
    rng = np.random.default_rng(5)
    forced = rng.integers(0, T, (20, 4 * nm)).astype(np.int32)
    n, res, cols, outl, thr = oracle.fit_subspace(traj, num_motions=nm, sigma=sigma, forced_cols=forced)
    got = ctx.fit_subspace(traj, num_motions=nm, sigma=sigma, forced_cols=forced)
    assert got["inliers"] == n and np.array_equal(got["best_cols"], cols) and np.array_equal(got["outlier"], outl)


def test_trajectories_into_fit_subspace(capi, oracle, seq640):
    # the live path of the node: calculateOpticalFlowTrajectory -> fitSubspace (node.cpp:295,348)
    frames, _ = seq640
    ctx = _ctx(capi, 640, 480, max_batch=4)
    tr = ctx.track_trajectories(frames[:5])
    full = tr["traj"][tr["len"] == 5]
    got = ctx.fit_subspace(full, num_motions=2, sigma=0.5, seed=3)
    n, res, cols, outl, thr = oracle.fit_subspace(full, num_motions=2, sigma=0.5, seed=3)
    assert got["inliers"] == n and np.array_equal(got["outlier"], outl)


# ---- VarFlow (VarFlow.cpp:600-697) ------------------------------------------------------------------------------
@pytest.mark.parametrize("size", [(96, 80), (320, 240), (333, 211)])
def test_varflow_matches_oracle(capi, oracle, size):
    w, h = size
    fr, _ = synth.sequence(w, h, 2, seed=5, camera=False, blobs=0, whole_field=(0.75, -0.5), margin=16)
    ctx = _ctx(capi, w, h)
    U, V = ctx.varflow(fr[0], fr[1])
    Uo, Vo = oracle.varflow(fr[0], fr[1])
    epe = np.sqrt((U - Uo) ** 2 + (V - Vo) ** 2)
    assert epe.mean() < FLOW_EPE_TOL and epe.max() < 1e-3
    # U is +x, V is y-UP: image motion (+0.75, -0.5)
    assert U.mean() > 0.3 and V.mean() > 0.2


def test_varflow_640_patch_sequence(capi, oracle):
    # C1: static camera, translating textured patch
    fr, _ = synth.sequence(640, 480, 2, seed=1234, camera=False, blobs=0, patch=True)
    ctx = _ctx(capi, 640, 480)
    U, V = ctx.varflow(fr[0], fr[1])
    Uo, Vo = oracle.varflow(fr[0], fr[1])
    epe = np.sqrt((U - Uo) ** 2 + (V - Vo) ** 2)
    assert epe.mean() < FLOW_EPE_TOL and epe.max() < 1e-3


def test_varflow_fast_mode_is_within_budget(capi, oracle):
    # eliding the (numerically inert) coarse-grid corrections, SURVEY 8a a14
    fr, _ = synth.sequence(320, 240, 2, seed=5, camera=False, blobs=0, whole_field=(0.75, -0.5), margin=16)
    ctx = _ctx(capi, 320, 240, vf_literal=0)
    U, V = ctx.varflow(fr[0], fr[1])
    Uo, Vo = oracle.varflow(fr[0], fr[1])
    assert np.sqrt((U - Uo) ** 2 + (V - Vo) ** 2).mean() < FLOW_EPE_TOL


def test_lk_points_per_cta_builds_are_equivalent():
    """k_lk_phase<NPTS> (points per one-warp CTA; MD_LK_NPTS selects the build, read once per process) must not change a bit:
    the request chain (next level of this point / first usable level of the next point) differs per build, the arithmetic
    does not.  Odd sizes: the grid is not a multiple of any NPTS and columns end inside a CTA."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, hashlib, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from motion_detection_b200 import capi, synth\n"
        "h = hashlib.sha256()\n"
        "for (w, hh, ps) in ((320, 240, 10), (333, 211, 7), (640, 480, 16)):\n"
        "    fr, _ = synth.sequence(w, hh, 4, seed=5)\n"
        "    ctx = capi.Context(width=w, height=hh, max_batch=3, pixel_step=ps, min_vector_size=0.2, seed=1)\n"
        "    r = ctx.process_batch(fr)\n"
        "    for k in ('next', 'status', 'keep', 'H', 'mask'):\n"
        "        h.update(np.ascontiguousarray(r[k]).tobytes())\n"
        "    ctx.close()\n"
        "print(h.hexdigest())\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    digests = {}
    for n in ("", "1", "2", "8", "16"):
        env = dict(os.environ)
        env.pop("MD_LK_NPTS", None)
        if n:
            env["MD_LK_NPTS"] = n
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        digests[n or "default"] = out.stdout.strip().splitlines()[-1]
    assert len(set(digests.values())) == 1, digests


def test_window_sum_kernels_are_equivalent():
    """The three ways the level records (and the Ix / Iy phase planes) of grid-mode LK are produced -- MD_WS_MODE 2 (default):
    k_window_sums_ring<true>, planes and window sums in one pass down the pyramid level; 1: k_phase_planes + the single-pass sums on
    the stored planes; 0: k_phase_planes + the sliding sums of k_window_sums -- are the same integers: whole batches must not change
    a bit.  pixel_step 5 / 8 / 10 / 20 / 40 cover every lattice step the ring kernel takes (5, 8, 10, 20, 40) and, at the levels where
    it does not (step 1, 2, 4), the mix of both kernels in one pyramid; odd sizes end lattice rows and columns inside a CTA."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, hashlib, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from motion_detection_b200 import capi, synth\n"
        "h = hashlib.sha256()\n"
        "for (w, hh, ps) in ((320, 240, 10), (333, 211, 5), (640, 480, 20), (641, 479, 8), (1920, 1080, 10), (800, 600, 40)):\n"
        "    fr, _ = synth.sequence(w, hh, 3, seed=11)\n"
        "    ctx = capi.Context(width=w, height=hh, max_batch=2, pixel_step=ps, min_vector_size=0.2, seed=1)\n"
        "    r = ctx.process_batch(fr)\n"
        "    for k in ('next', 'status', 'keep', 'H', 'mask'):\n"
        "        h.update(np.ascontiguousarray(r[k]).tobytes())\n"
        "    ctx.close()\n"
        "print(h.hexdigest())\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    digests = {}
    for m in ("", "1", "0", "rows320", "rows60"):
        env = dict(os.environ)
        env.pop("MD_WS_MODE", None)
        env.pop("MD_WS_ROWS", None)
        if m.startswith("rows"):
            env["MD_WS_ROWS"] = m[4:]            # plane rows per CTA segment of the single-pass kernel (launch geometry only)
        elif m:
            env["MD_WS_MODE"] = m
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        digests[m or "default"] = out.stdout.strip().splitlines()[-1]
    assert len(set(digests.values())) == 1, digests
