"""GPU edge cases through the C ABI: tiny / ragged sizes, narrow images (the warp's column-block rule changes below
64 columns), 4K with projective terms, generic LK window, argument errors, batch limits."""
import numpy as np
import pytest

from motion_detection_b200 import synth

pytestmark = pytest.mark.gpu


def _ctx(capi, w, h, **kw):
    kw.setdefault("min_vector_size", 0.2)
    return capi.Context(width=w, height=h, **kw)


@pytest.mark.parametrize("size", [(48, 40), (63, 17), (8, 8), (100, 12), (70, 65)])
def test_mask_small_and_narrow_images(capi, oracle, size):
    # w < 64: WarpPerspectiveInvoker's column block is the whole row (bw0 = w); h < 16 changes bh0 as well
    w, h = size
    rng = np.random.default_rng(w * 100 + h)
    a = rng.integers(0, 256, (h, w), dtype=np.uint8)
    b = rng.integers(0, 256, (h, w), dtype=np.uint8)
    ctx = _ctx(capi, w, h)
    for H in (np.eye(3), np.array([[1.02, 0.01, -1.3], [-0.02, 0.97, 0.6], [3e-4, -2e-4, 1.0]]),
              np.array([[1, 0, 0.015625], [0, 1, -0.484375], [0, 0, 1.0]])):
        for thresh in (190, 50):
            assert np.array_equal(ctx.motion_mask(a, b, H, thresh=thresh), oracle.motion_mask(a, b, H, thresh=thresh)), (size, thresh)


def test_singular_homography_is_handled(capi, oracle):
    w, h = 64, 48
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (h, w), dtype=np.uint8)
    b = rng.integers(0, 256, (h, w), dtype=np.uint8)
    ctx = _ctx(capi, w, h)
    H = np.array([[1, 2, 3], [2, 4, 6], [0, 0, 1.0]])            # det = 0: cv::invert returns the zero matrix
    assert np.array_equal(ctx.motion_mask(a, b, H, thresh=60), oracle.motion_mask(a, b, H, thresh=60))


def test_single_level_pyramid_and_tiny_flow(capi, oracle):
    # 100 x 90: buildOpticalFlowPyramid keeps two levels (50 x 45 > 40), 80 x 60 keeps one
    for (w, h, levels) in ((100, 90, 2), (80, 60, 1)):
        fr, _ = synth.sequence(w, h, 2, seed=3, blobs=1, margin=32)
        ctx = _ctx(capi, w, h, pixel_step=7)
        assert ctx.levels == levels == len(oracle.pyramid(fr[0]))
        ctx.pyramid(fr[0], 0)
        ctx.pyramid(fr[1], 1)
        nxt, st = ctx.lk_flow(0, 1)
        ref, rst = oracle.lk(fr[0], fr[1], ctx.grid_points())
        assert (st != rst).mean() < 0.02
        ok = (st == 1) & (rst == 1)
        if ok.any():
            assert np.linalg.norm(nxt[ok] - ref[ok], axis=1).mean() < 0.01


def test_generic_window_kernel(capi, oracle):
    # windows other than the reference's 40 run the generic kernel (k_lk)
    fr, _ = synth.sequence(320, 240, 2, seed=9, blobs=1)
    for win in (21, 32):
        ctx = _ctx(capi, 320, 240, lk_win=win, pixel_step=12)
        ctx.pyramid(fr[0], 0)
        ctx.pyramid(fr[1], 1)
        nxt, st = ctx.lk_flow(0, 1)
        ref, rst = oracle.lk(fr[0], fr[1], ctx.grid_points(), win=win)
        assert (st != rst).mean() < 0.01
        ok = (st == 1) & (rst == 1)
        assert np.linalg.norm(nxt[ok] - ref[ok], axis=1).mean() < 0.01


def test_4k_projective_chain(capi, oracle):
    # C3: 3840 x 2160, homography = affine + (h31, h32) = (1e-6, -2e-6); 6 pyramid levels
    w, h = 3840, 2160
    fr, Hs = synth.sequence(w, h, 2, seed=1234, h31=1e-6, h32=-2e-6)
    ctx = _ctx(capi, w, h, pixel_step=40, seed=5)
    assert ctx.levels == 6
    res = ctx.process_batch(fr)
    ref = oracle.process_pair(fr[0], fr[1], pixel_step=40, min_vector_size=0.2, seed=5)
    ok = (res["status"][0] == 1) & (ref["status"] == 1)
    assert (res["status"][0] != ref["status"]).mean() < 0.005
    assert np.linalg.norm(res["next"][0][ok] - ref["next"][ok], axis=1).mean() < 0.01
    assert np.linalg.norm(res["H"][0] - ref["H"]) / np.linalg.norm(ref["H"]) < 1e-4
    assert (res["mask"][0] == ref["mask"]).mean() >= 0.999
    assert np.linalg.norm(res["H"][0] - Hs[0]) / np.linalg.norm(Hs[0]) < 0.02


def test_argument_errors_are_status_codes(capi):
    import ctypes as C
    ctx = _ctx(capi, 160, 120, max_batch=2)
    frames = np.zeros((5, 120, 160), np.uint8)
    with pytest.raises(capi.MotionB200Error) as e:
        ctx.process_batch(frames)                       # 4 pairs > max_batch
    assert e.value.code == -1
    with pytest.raises(capi.MotionB200Error) as e:
        ctx.lk_flow(0, 7)                               # slot out of range
    assert e.value.code == -1
    with pytest.raises(capi.MotionB200Error):
        ctx.fit_subspace(np.zeros((10, 20, 2), np.float32))      # 2F > 32
    lib = capi.lib()
    assert lib.md_process_batch(ctx._h, None, None, 0) == -1
    # nothing was written on error and the context is still usable
    fr, _ = synth.sequence(160, 120, 3, seed=2, blobs=1, margin=32)
    res = ctx.process_batch(fr)
    assert res["mask"].shape == (2, 120, 160)


def test_static_scene_min_vector_filter(capi, oracle):
    # C1: static camera; with min_vector_size = 0.2 only the moving patch survives the filter (cpp:86)
    fr, _ = synth.sequence(640, 480, 2, seed=1234, camera=False, blobs=0, patch=True)
    ctx = _ctx(capi, 640, 480, seed=4)
    res = ctx.process_batch(fr)
    ref = oracle.process_pair(fr[0], fr[1], min_vector_size=0.2, seed=4)
    assert abs(int(res["num_vectors"][0]) - ref["num_vectors"]) <= 3
    assert np.array_equal(res["keep"][0] != 0, ref["keep"] != 0) or (res["keep"][0] != ref["keep"]).mean() < 0.002
    assert (res["mask"][0] == ref["mask"]).mean() >= 0.999


def test_huge_pixel_step_falls_back_to_point_kernel(capi, oracle):
    """pixel_step larger than the window-sums block: the grid mode must quietly use the per-point kernel (not fail)."""
    import numpy as np
    from motion_detection_b200 import synth
    w, h = 1280, 720
    frames, _ = synth.sequence(w, h, 2, seed=4, blobs=1)
    ctx = capi.Context(width=w, height=h, max_batch=1, pixel_step=500, min_vector_size=0.2)
    ctx.pyramid(frames[0], 0)
    ctx.pyramid(frames[1], 1)
    nxt, st = ctx.lk_flow(0, 1)
    ref, rst = oracle.lk(frames[0], frames[1], ctx.grid_points())
    assert np.array_equal(st, rst)
    assert np.abs(nxt[st == 1] - ref[rst == 1]).max() < 1e-2


def test_lk_large_displacement_restages_the_next_frame_tile(capi, oracle):
    """A field moving 13 x 9 px per frame: at every pyramid level the window drifts out of the staged next-frame tile
    (margins 8 x 3 px), so the TMA restage path runs in both LK kernels; results must still be the oracle's."""
    import numpy as np
    from motion_detection_b200 import synth
    w, h = 640, 480
    frames, _ = synth.sequence(w, h, 2, seed=8, camera=False, blobs=0, whole_field=(13.0, -9.0))
    ctx = capi.Context(width=w, height=h, max_batch=1, pixel_step=10, min_vector_size=0.2)
    ctx.pyramid(frames[0], 0)
    ctx.pyramid(frames[1], 1)
    pts = ctx.grid_points()
    a, sa = ctx.lk_flow(0, 1)                    # phase-plane kernel
    b, sb = ctx.lk_flow(0, 1, pts)               # per-point kernel
    assert np.array_equal(sa, sb) and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    ref, rst = oracle.lk(frames[0], frames[1], pts)
    assert (sa != rst).mean() < 0.005
    ok = (sa == 1) & (rst == 1)
    assert np.linalg.norm(a[ok] - ref[ok], axis=1).mean() < 0.01
    inner = ok & (pts[:, 0] > 80) & (pts[:, 0] < w - 80) & (pts[:, 1] > 80) & (pts[:, 1] < h - 80)
    assert np.abs(np.median(a[inner] - pts[inner], axis=0) - np.array([13.0, -9.0])).max() < 0.1


# ---- K4 fast path (TMA-staged boxes + fixed-point coordinates): bit-exactness over many near-identity homographies ---------------
@pytest.mark.parametrize("size", [(1920, 1080), (641, 479), (3840, 2160)])
def test_mask_fast_path_random_homographies(capi, oracle, size):
    """Frame-to-frame egomotion is near-identity: those tiles run from TMA boxes with the expansion-based reciprocal and the
    2^-15 rounding guard.  Random small rotations / zooms / translations / projective terms, sub-pixel translations that put
    EVERY pixel on a rounding boundary (multiples of 1/64 px), and homographies strong enough to send some tiles to the gather
    path: the mask must equal the oracle's bit for bit (warpPerspective + absdiff + threshold + erode + dilate,
    optical_flow_calculator.cpp:124-127, background_subtractor.cpp:31-32)."""
    w, h = size
    rng = np.random.default_rng(w + h)
    prev = rng.integers(0, 256, (h, w), dtype=np.uint8)
    cur = rng.integers(0, 256, (h, w), dtype=np.uint8)
    ctx = _ctx(capi, w, h)
    cases = []
    n_rand = 6 if w < 3000 else 2
    for _ in range(n_rand):
        a = rng.normal(0, 0.002)
        s = 1.0 + rng.normal(0, 0.002)
        H = np.array([[s * np.cos(a), -s * np.sin(a), rng.normal(0, 3.0)], [s * np.sin(a), s * np.cos(a), rng.normal(0, 3.0)],
                      [rng.normal(0, 2e-6), rng.normal(0, 2e-6), 1.0]])
        cases.append(H)
    cases.append(np.array([[1, 0, 3 / 64.0], [0, 1, -5 / 64.0], [0, 0, 1.0]]))          # every coordinate is a rounding tie
    cases.append(np.array([[1, 0, 7.0], [0, 1, -2.0], [0, 0, 1.0]]))                    # integer shift: fractions are all 0
    cases.append(np.array([[1.0, 0, 0.25], [0, 1.0, 0.75], [3e-5, -2e-5, 1.0]]))        # projective terms near the expansion limit
    cases.append(np.array([[1.12, 0.05, -40.0], [-0.04, 1.1, 12.0], [0, 0, 1.0]]))      # zoom: boxes near / over the size limit
    cases.append(np.array([[np.cos(0.2), -np.sin(0.2), 100.0], [np.sin(0.2), np.cos(0.2), -80.0], [0, 0, 1.0]]))   # rotation: gather tiles
    cases.append(np.array([[1, 0, -float(w)], [0, 1, 0], [0, 0, 1.0]]))                 # the source lies entirely outside the image
    for H in cases:
        for thresh in (190, 60):
            got = ctx.motion_mask(prev, cur, H, thresh=thresh, morph=True)
            ref = oracle.motion_mask(prev, cur, H, thresh=thresh, morph=True)
            assert np.array_equal(got, ref), (size, H.tolist(), thresh, int((got != ref).sum()))
        got = ctx.motion_mask(prev, cur, H, thresh=60, morph=False)
        assert np.array_equal(got, oracle.motion_mask(prev, cur, H, thresh=60, morph=False)), (size, H.tolist(), "raw")


@pytest.mark.parametrize("size", [(320, 240), (641, 479), (1920, 1080)])
def test_packed_mask_output_equals_byte_mask(capi, size):
    """md_config.mask_packed: the same mask as 1 bit per pixel (LSB first), an eighth of the bytes on the way back to the host."""
    w, h = size
    frames, _ = synth.sequence(w, h, 4, seed=31, blobs=3)
    a = _ctx(capi, w, h, max_batch=3, seed=5, diff_threshold=40)
    b = _ctx(capi, w, h, max_batch=3, seed=5, diff_threshold=40, mask_packed=1)
    ra, rb = a.process_batch(frames), b.process_batch(frames)
    assert ra["mask"].any()
    assert rb["mask_bits"].shape == (3, h, (w + 7) // 8)
    assert np.array_equal(ra["mask"], rb["mask"])
    assert a.stats()["mask_pixels"] == b.stats()["mask_pixels"] == int((ra["mask"] > 0).sum())
    for k in ("next", "status", "H", "inliers"):
        assert np.array_equal(ra[k], rb[k]), k
