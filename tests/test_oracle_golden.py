"""CPU: the plain-C oracle against the committed golden vectors (tests/golden/golden_cv2.npz, produced from the OpenCV
routines the reference calls by tests/golden/make_golden.py).  Integer stages bit-exact; LK and VarFlow to float noise."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_cv2.npz"))


def test_gray(oracle):
    assert np.array_equal(oracle.gray(G["gray_in"]), G["gray_out"])


def test_pyramid_and_scharr(oracle):
    pyr = oracle.pyramid(G["pyr_in"])
    assert len(pyr) == int(G["pyr_levels"])
    for l, im in enumerate(pyr):
        assert np.array_equal(im, G["pyr_l%d" % l])
        assert np.array_equal(oracle.scharr(im), G["pyr_d%d" % l])


def test_pyramid_level_rule(oracle):
    # cv::buildOpticalFlowPyramid stops when the next level would be <= winSize in either dimension
    assert oracle.pyr_levels(320, 240) == 2 and oracle.pyr_levels(640, 480) == 3
    assert oracle.pyr_levels(1920, 1080) == 4 and oracle.pyr_levels(3840, 2160) == 5
    assert oracle.pyr_levels(120, 100) == 1 and oracle.pyr_levels(60, 60) == 0


def test_lk(oracle):
    nxt, st = oracle.lk(G["lk_f0"], G["lk_f1"], G["lk_pts"])
    assert (st != G["lk_status"]).mean() < 0.01
    ok = (st == 1) & (G["lk_status"] == 1)
    d = np.linalg.norm(nxt[ok] - G["lk_next"][ok], axis=1)
    assert d.mean() < 1e-3 and np.median(d) < 1e-4      # north_star flow bar is 0.01 px mean EPE


def test_perspective_4pt(oracle):
    ok, H = oracle.perspective_4pt(G["p4_src"], G["p4_dst"])
    assert ok == 1
    assert np.linalg.norm(H - G["p4_H"]) / np.linalg.norm(G["p4_H"]) < 1e-9


def test_degenerate_first_four_is_reported(oracle):
    # the reference's literal first-four choice is collinear (column x = 0): no valid homography
    src = np.array([[0, 0], [0, 10], [0, 20], [0, 30]], np.float64)
    ok, _ = oracle.perspective_4pt(src, src + 1.5)
    assert ok == 0


def test_warp_and_mask_chain(oracle):
    a, b = G["mask_a"], G["mask_b"]
    for i, H in enumerate(G["mask_H"]):
        assert np.array_equal(oracle.warp_perspective(a, H), G["mask_warp"][i])
        assert np.array_equal(oracle.motion_mask(a, b, H, thresh=40), G["mask_t40"][i])
        assert np.array_equal(oracle.motion_mask(a, b, H, thresh=40, morph=False), G["mask_t40_nomorph"][i])
    assert (G["mask_t40"] > 0).sum() > 500


def test_threshold_is_strict(oracle):
    a = np.full((8, 8), 190, np.uint8)
    z = np.zeros((8, 8), np.uint8)
    assert oracle.absdiff_threshold(a, z, 190).sum() == 0
    assert (oracle.absdiff_threshold(a + 1, z, 190) == 255).all()


def test_morphology_border_identities(oracle):
    full = np.full((5, 7), 255, np.uint8)
    assert (oracle.erode3(full) == 255).all()            # out-of-image pixels do not erode
    one = np.zeros((5, 7), np.uint8); one[0, 0] = 255
    d = oracle.dilate3(one)
    assert d[:2, :2].min() == 255 and d.sum() == 4 * 255


def test_varflow(oracle):
    U, V = oracle.varflow(G["vf_a"], G["vf_b"])
    assert np.abs(U - G["vf_U"]).max() < 2e-5 and np.abs(V - G["vf_V"]).max() < 2e-5
    # U is +x, V is y-UP (VarFlow.cpp:103-107): the field moves by (+0.75, -0.5) px in image coordinates
    assert U.mean() > 0.3 and V.mean() > 0.2


def test_varflow_primitives(oracle):
    x = G["prim_in"]
    assert np.abs(oracle.gaussian_blur_f32(x, 1.5) - G["prim_blur15"]).max() < 5e-5
    assert np.abs(oracle.gaussian_blur_f32(x, 2.8) - G["prim_blur28"]).max() < 5e-5
    assert np.abs(oracle.resize_linear_f32(x, 23, 16) - G["prim_resize_23_16"]).max() < 1e-3
    assert np.abs(oracle.resize_linear_f32(x, 94, 66) - G["prim_resize_94_66"]).max() < 1e-3
    assert np.abs(oracle.resize_linear_f32(G["prim_in2"], 24, 16) - G["prim_resize_half"]).max() < 1e-5


def test_glibc_rand_known_answers(oracle):
    # glibc TYPE_3 rand(): srand(1) (also the default state) starts 1804289383, 846930886, 1681692777, ...
    assert oracle.glibc_rand(1, 5) == [1804289383, 846930886, 1681692777, 1714636915, 1957747793]
    assert oracle.glibc_rand(42, 3) == [71876166, 708592740, 1483128881]


def test_glibc_rand_matches_libc(oracle):
    import ctypes
    try:
        libc = ctypes.CDLL("libc.so.6")
    except OSError:
        pytest.skip("no glibc")
    for seed in (1, 7, 123456789):
        libc.srand(seed)
        assert [libc.rand() for _ in range(400)] == oracle.glibc_rand(seed, 400)


# ---- live path restatements (oracle/md_oracle_live.c): hand-worked cases ------------------------------------------------
def test_cluster_euclidean_hand_worked():
    from oracle import oracle as O
    # points arrive in this order; threshold 12: (0,0) founds 0; (10,0) joins 0 (d = 10); (0,10) joins 0; (100,100) founds 1;
    # (10,10) is 10 away from two members of cluster 0; (105,100) joins 1; (5,5) joins 0
    pts = np.array([[0, 0], [10, 0], [0, 10], [100, 100], [10, 10], [105, 100], [5, 5]], np.float32)
    labels, ncl, boxes, sizes, ids = O.cluster_euclidean(pts, 12.0, 1)
    assert labels.tolist() == [0, 0, 0, 1, 0, 1, 0] and ncl == 2
    # cv::boundingRect of the rounded points: tl = min, br = max + 1 (the CSV row of MotionLogger::writeBoundingBox)
    assert boxes.tolist() == [[0, 0, 11, 11], [100, 100, 106, 101]] and sizes.tolist() == [5, 2] and ids.tolist() == [0, 1]
    # only clusters with MORE than 5 members survive with the reference's constant (flow_clusterer.cpp:264)
    assert len(O.cluster_euclidean(pts, 12.0, 5)[2]) == 0
    # a tie between an older and a younger cluster goes to the older one (strict '<' while scanning in creation order)
    pts = np.array([[0, 0], [20, 0], [10, 0]], np.float32)
    assert O.cluster_euclidean(pts, 11.0, 0)[0].tolist() == [0, 1, 0]
    # order dependence: the same points in another order chain into one cluster
    pts = np.array([[0, 0], [10, 0], [20, 0]], np.float32)
    assert O.cluster_euclidean(pts, 11.0, 0)[0].tolist() == [0, 0, 0]
    # rounding is half-to-even, like saturate_cast<int>(float)
    pts = np.array([[0.5, 1.5], [2.5, 3.5]], np.float32)
    assert O.cluster_euclidean(pts, 10.0, 0)[2].tolist() == [[0, 2, 3, 5]]


def test_traj_step_bookkeeping():
    from oracle import oracle as O
    import ctypes as C
    P, F, w, h = 4, 3, 100, 80
    cur = np.array([[20, 20], [30, 30], [40, 40], [50, 50]], np.float32)
    traj = np.zeros((P, F, 2), np.float32); traj[:, 0] = cur
    ln = np.ones(P, np.int32)
    nxt = np.array([[21, 21], [5, 30], [41, 75], [51, 51]], np.float32)       # 2nd leaves the margin in x, 3rd in y (>= h - 10)
    st = np.array([1, 1, 1, 0], np.uint8)
    O.lib().orc_traj_step(cur.ctypes.data_as(O.f32p), nxt.ctypes.data_as(O.f32p), st.ctypes.data_as(O.u8p),
                          traj.ctypes.data_as(O.f32p), ln.ctypes.data_as(O.i32p), P, F, w, h)
    assert ln.tolist() == [2, 1, 1, 1]
    assert cur.tolist() == [[21, 21], [30, 30], [40, 40], [50, 50]]
    assert traj[0, 1].tolist() == [21, 21]


def test_find_outliers_against_numpy_statistics():
    """orc_find_outliers is the literal createMask loop; numpy's median / MAD give the same statistics and flags."""
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    for n, incl in [(500, False), (501, True), (64, False), (1, False)]:
        d = rng.normal([1.2, -0.8], 0.1, (n, 2)).astype(np.float32).astype(np.float64)
        d[rng.random(n) < 0.2] = 0.0                                   # filtered / failed vectors are stored as zeros
        d[rng.random(n) < 0.05] += 3.0                                 # movers
        out, st = O.find_outliers(d, incl)
        sel = np.ones(n, bool) if incl else (np.abs(d) > 0).any(1)
        if not sel.any():
            assert out.sum() == 0
            continue
        ang, mag = np.arctan2(d[:, 1], d[:, 0]), np.sqrt(d[:, 1] ** 2 + d[:, 0] ** 2)
        flags = np.zeros(n, bool)
        for q, v in enumerate((ang, mag)):
            med = np.median(v[sel]); mad = np.median(np.abs(v[sel] - med))
            assert st[2 * q] == med and st[2 * q + 1] == mad
            with np.errstate(divide="ignore", invalid="ignore"):
                flags |= sel & (np.abs(0.6745 * np.abs(v - med) / mad) > 3.5)
        assert np.array_equal(out.astype(bool), flags)
    # no participating vector: nothing is flagged (vals.empty() return, outlier_detector.cpp:142-145)
    out, st = O.find_outliers(np.zeros((10, 2)), False)
    assert out.sum() == 0


def test_line_aa_golden(oracle):
    """cv::line(..., 1, CV_AA) -- OpenCV's LineAA restated in oracle/md_oracle_draw.c -- against frozen cv2 output: 40 segments
    drawn one after another (clipped ones included), 3 channels and 1 channel."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_cv2.npz"))
    img3, img1 = g["aa_img3"].copy(), g["aa_img1"].copy()
    for s, c in zip(g["aa_segs"], g["aa_cols"]):
        oracle.line_aa(img3, (s[0], s[1]), (s[2], s[3]), c)
        oracle.line_aa(img1, (s[0], s[1]), (s[2], s[3]), c[:1])
    assert np.array_equal(img3, g["aa_out3"])
    assert np.array_equal(img1, g["aa_out1"])


def test_show_optical_flow_vectors_golden(oracle):
    """OpticalFlowVisualizer::showOpticalFlowVectors (optical_flow_visualizer.cpp:23-71): arrows of a 12 x 9 field, overlapping."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_cv2.npz"))
    out, n = oracle.draw_flow(g["flowdraw_img"], g["flowdraw_vec"], 10, 0.2, (255, 0, 0))
    assert n == int(g["flowdraw_n"]) and n > 50
    assert np.array_equal(out, g["flowdraw_out"])
