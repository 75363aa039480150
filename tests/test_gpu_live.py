"""GPU parity of the node's live path (MotionDetectionNode::imageCallback): ring of pyramids -> trajectories -> fitSubspace ->
outlier points -> clusterEuclidean -> bounding boxes, through the C ABI, against the CPU oracle's literal restatement."""
import numpy as np
import pytest

from motion_detection_b200 import synth

pytestmark = pytest.mark.gpu


def _ctx(capi, w, h, **kw):
    kw.setdefault("min_vector_size", 0.2)
    return capi.Context(width=w, height=h, **kw)


@pytest.mark.parametrize("n,thr,layout", [(400, 50.0, "blobs"), (1500, 12.0, "uniform"), (900, 10.0, "grid"), (1200, 15.0, "grid"),
                                          (7, 12.0, "tiny"), (1, 5.0, "tiny"), (3000, 3.0, "dupes")])
def test_cluster_points_identical_to_oracle(capi, oracle, n, thr, layout):
    """clusterEuclidean is greedy and order dependent; the device version (nearest EARLIER neighbour + sequential label pass)
    must give the same labels, the same cluster order and the same boxes -- including exact distance ties (grid layout:
    many equidistant neighbours in different clusters) and duplicate points."""
    rng = np.random.default_rng(n)
    if layout == "blobs":
        c = rng.uniform(50, 600, (6, 2))
        pts = c[rng.integers(0, 6, n)] + rng.normal(0, 12, (n, 2))
    elif layout == "uniform":
        pts = rng.uniform(0, 640, (n, 2))
    elif layout == "grid":
        g = np.stack(np.meshgrid(np.arange(0, 400, 10), np.arange(0, 300, 10), indexing="ij"), -1).reshape(-1, 2)
        pts = g[rng.permutation(len(g))[:n]].astype(np.float64)
    elif layout == "dupes":
        pts = rng.integers(0, 40, (n, 2)).astype(np.float64) * 2.5
    else:
        pts = np.array([[0, 0], [10, 0], [0, 10], [100, 100], [10, 10], [105, 100], [5, 5]], np.float64)[:n]
    pts = pts.astype(np.float32)
    ctx = _ctx(capi, 640, 480)
    for min_size in (5, 0):
        ref = oracle.cluster_euclidean(pts, thr, min_size)
        got = ctx.cluster_points(pts, thr, min_size)
        assert got[1] == ref[1]
        assert np.array_equal(got[0], ref[0])
        for a, b in zip(got[2:], ref[2:]):
            assert np.array_equal(a, b)


def test_cluster_points_empty(capi):
    ctx = _ctx(capi, 640, 480)
    labels, nall, boxes, sizes, ids = ctx.cluster_points(np.zeros((0, 2), np.float32))
    assert len(labels) == 0 and nall == 0 and len(boxes) == 0


def _moving_patch_sequence(w, h, F, seed):
    """textured background (static camera) + a textured patch translating 3 px / frame: its trajectories leave the motion
    subspace of the background."""
    frames, _ = synth.sequence(w, h, F, seed=seed, blobs=0)
    frames = np.repeat(frames[:1], F, axis=0).copy()
    rng = np.random.default_rng(seed)
    patch = rng.integers(120, 255, (90, 120), dtype=np.uint8)
    import scipy.ndimage as ndi
    patch = ndi.gaussian_filter(patch.astype(np.float32), 1.5).astype(np.uint8)
    for f in range(F):
        x0, y0 = 200 + 3 * f, 150 + 2 * f
        frames[f, y0:y0 + 90, x0:x0 + 120] = patch
    return frames


@pytest.mark.parametrize("nm,sigma", [(2, 0.5), (1, 1.0)])
def test_window_detect_matches_oracle(capi, oracle, nm, sigma):
    w, h = 640, 480
    F = 2 * nm + 1
    frames = _moving_patch_sequence(w, h, F + 2, seed=5)
    ctx = _ctx(capi, w, h, max_batch=F - 1)
    for k in range(F + 2):
        fill = ctx.window_push(frames[k])
        assert fill == min(k + 1, F)
        if fill < F:
            with pytest.raises(capi.MotionB200Error):
                ctx.window_detect(num_motions=nm)
            continue
        got = ctx.window_detect(num_motions=nm, sigma=sigma, distance_threshold=50.0, seed=3 + k)
        win = frames[k - F + 1:k + 1]
        ref = oracle.live_detect(win, num_motions=nm, sigma=sigma, distance_threshold=50.0, seed=3 + k)
        # trajectories: same set of complete trajectories (LK differs by ~1e-5 px, so allow a few border flips)
        common, gi, ri = np.intersect1d(got["traj_index"], ref["traj_index"], return_indices=True)
        assert len(common) >= 0.995 * max(len(ref["traj_index"]), 1)
        d = np.linalg.norm(got["traj"][gi] - ref["traj"][ri], axis=2)
        assert d.mean() < 0.01
        # the subspace fit on the device's OWN trajectories must equal the oracle run on those same trajectories
        n2, res2, cols2, outl2, _ = oracle.fit_subspace(got["traj"], num_motions=nm, sigma=sigma, seed=3 + k)
        assert got["subspace_inliers"] == n2
        assert np.array_equal(got["best_cols"], cols2)
        assert np.array_equal(got["outlier"], outl2)
        opts = got["traj"][outl2 != 0][:, F - 2]
        assert np.array_equal(got["outlier_points"], opts)
        lab, nall, boxes, sizes, ids = oracle.cluster_euclidean(opts, 50.0, 5)
        assert got["num_clusters_all"] == nall and np.array_equal(got["labels"], lab)
        assert np.array_equal(got["boxes"], boxes) and np.array_equal(got["cluster_sizes"], sizes)
        assert np.array_equal(got["cluster_ids"], ids)
        # end to end against the oracle's own chain (its trajectories differ by ~1e-5 px, so residuals that sit on the
        # threshold may flip): the outlier flags of the common trajectories agree
        agree = (got["outlier"][gi] == ref["outlier"][ri]).mean()
        assert agree >= 0.97, agree


def test_window_ring_equals_fresh_tracking(capi):
    """The ring reuses pyramids across callbacks; the result must equal md_track_trajectories on the same F frames."""
    w, h, F = 320, 240, 5
    frames, _ = synth.sequence(w, h, F + 3, seed=77)
    ctx = _ctx(capi, w, h, max_batch=F - 1)
    fresh = _ctx(capi, w, h, max_batch=F - 1)
    for k in range(F + 3):
        fill = ctx.window_push(frames[k])
        if fill < F:
            continue
        got = ctx.window_detect(num_motions=2, sigma=0.5)
        ref = fresh.track_trajectories(frames[k - F + 1:k + 1])
        idx = np.nonzero(ref["len"] == F)[0]
        assert np.array_equal(got["traj_index"], idx)
        assert np.array_equal(got["traj"], ref["traj"][idx])


def test_live_node_mirrors_image_callback(capi, oracle):
    """LiveNode = imageCallback without ROS: skip_frames, window fill, frame counters and the CSV rows of the logger."""
    from motion_detection_b200.node import LiveNode
    w, h, nm, skip = 320, 240, 1, 2
    F = 2 * nm + 1
    frames, _ = synth.sequence(w, h, 11, seed=9, blobs=2)
    rgb = np.repeat(frames[..., None], 3, axis=3)                  # the node converts to "rgb8" (node.cpp:271)
    node = LiveNode(w, h, num_motions=nm, skip_frames=skip, sigma=1.0, distance_threshold=50.0, seed=4)
    kept, ncb = [], 0
    for g in range(len(frames)):
        res = node.on_image(rgb[g])
        if g % skip:
            assert res is None
            continue
        kept.append(g)
        if len(kept) < F:
            assert res is None
            continue
        assert res is not None and res["frame"] == g
        win = [frames[i] for i in kept[-F:]]                       # weights sum to 2^15: gray(rgb replicated) == gray
        ref = oracle.live_detect(win, num_motions=nm, sigma=1.0, distance_threshold=50.0, seed=4 + ncb)
        ncb += 1
        assert abs(res["num_trajectories"] - ref["num_trajectories"]) <= 0.005 * ref["num_trajectories"] + 1
        n2, _, cols2, outl2, _ = oracle.fit_subspace(res["traj"], num_motions=nm, sigma=1.0, seed=3 + ncb)
        assert np.array_equal(res["outlier"], outl2)
    assert node.callbacks == ncb and node.frame_number == ncb and node.global_frame_count == len(frames)
    for row in node.log_rows:
        f = [int(v) for v in row.split(", ")]
        assert len(f) == 6 and f[0] in kept and f[4] > f[2] and f[5] > f[3]
    node.close()


@pytest.mark.parametrize("n,incl", [(20736, False), (20736, True), (3072, False), (501, True), (2, False), (1, True), (300000, False)])
def test_find_outliers_matches_oracle(capi, oracle, n, incl):
    """findOutliers / createMask: exact medians by radix select.  Magnitudes are bit-exact; angles go through atan2 (device
    libm vs glibc, <= 2 ulp), so the angle statistics are compared to 1e-14 relative and the flags must agree except where a
    z-score sits within 1e-9 of the 3.5 threshold."""
    rng = np.random.default_rng(n)
    d = rng.normal([1.2, -0.8], 0.15, (n, 2)).astype(np.float32).astype(np.float64)
    d[rng.random(n) < 0.2] = 0.0
    d[rng.random(n) < 0.03] += rng.normal(0, 2.0, 2)
    ctx = _ctx(capi, 640, 480)
    got, gst = ctx.find_outliers(d, incl)
    ref, rst = oracle.find_outliers(d, incl)
    assert gst[2] == rst[2] and gst[3] == rst[3]                       # magnitude median / MAD: bit-exact
    assert np.allclose(gst[:2], rst[:2], rtol=1e-14, atol=1e-15)
    diff = np.nonzero(got != ref)[0]
    if len(diff):
        sel = np.ones(n, bool) if incl else (np.abs(d) > 0).any(1)
        ang = np.arctan2(d[:, 1], d[:, 0])
        z = 0.6745 * np.abs(ang - rst[0]) / rst[1]
        assert np.all(np.abs(z[diff] - 3.5) < 1e-9) and sel[diff].all()


def test_find_outliers_degenerate(capi, oracle):
    ctx = _ctx(capi, 640, 480)
    z = np.zeros((100, 2))
    got, st = ctx.find_outliers(z, False)                              # no participating vector
    assert got.sum() == 0 and np.all(st == 0)
    got, st = ctx.find_outliers(z, True)                               # MAD = 0: 0 / 0 is not > 3.5
    ref, rst = oracle.find_outliers(z, True)
    assert np.array_equal(got, ref) and got.sum() == 0
    d = np.tile([[1.0, 0.0]], (50, 1)); d[7] = [5.0, 0.0]             # MAD = 0 and one mover: x / 0 = inf > 3.5
    got, st = ctx.find_outliers(d, False)
    ref, rst = oracle.find_outliers(d, False)
    assert np.array_equal(got, ref) and got[7] == 1 and got.sum() == 1


# ---- clusterEuclidean: f32-rounded distances (point_cluster.cpp:62-65, sqrt(float) overload) --------------------------------
def test_cluster_points_near_threshold_and_collapsed_ties(capi, oracle):
    """(i) a squared distance whose f64 root is below the threshold but whose f32 root equals it: the reference (strict '<'
    on the f32 root) founds a new cluster; (ii) two different squared distances that collapse to the same f32 root: the
    reference keeps the OLDER cluster."""
    ctx = _ctx(capi, 640, 480)
    # (i) d2 = nextafter(2500, 0) in f32: sqrt -> 49.999998 in f64 but 50.0f in f32
    d2 = np.nextafter(np.float32(2500.0), np.float32(0.0))
    x = np.sqrt(np.float64(d2))
    pts = np.array([[0.0, 0.0], [x, 0.0]], np.float32)
    got_d2 = np.float32(pts[1, 0]) * np.float32(pts[1, 0])
    ref = oracle.cluster_euclidean(pts, 50.0, 0)
    got = ctx.cluster_points(pts, 50.0, 0)
    assert np.array_equal(got[0], ref[0]) and got[1] == ref[1]
    if np.sqrt(np.float32(got_d2), dtype=np.float32) == np.float32(50.0) and np.sqrt(np.float64(got_d2)) < 50.0:
        assert ref[1] == 2                     # not joined although the exact root is below the threshold
    # (ii) search squares near 1e6 for two distinct f32 values with the same f32 root
    base = np.float32(1.0e6)
    cands = [base]
    for _ in range(6):
        cands.append(np.nextafter(cands[-1], np.float32(2e6)))
    roots = [np.sqrt(c, dtype=np.float32) for c in cands]
    pair = next((i for i in range(len(cands) - 1) if roots[i] == roots[i + 1]), None)
    assert pair is not None
    # point 2 sits at distances sqrt(cands[pair + 1]) from point 0 (cluster 0) and sqrt(cands[pair]) from point 1 (cluster 1):
    # the younger cluster is nearer in d2, equal in f32 distance -> the reference keeps cluster 0
    rng = np.random.default_rng(1)
    for trial in range(200):
        a = np.float32(rng.uniform(900, 1100))
        pts = np.array([[0, 0], [3000, 0], [a, 0]], np.float32)
        ref = oracle.cluster_euclidean(pts, 5000.0, 0)
        got = ctx.cluster_points(pts, 5000.0, 0)
        assert np.array_equal(got[0], ref[0]) and got[1] == ref[1]
    # dense random layouts with f32 coordinates on a coarse lattice produce many collapsed ties
    for seed in range(4):
        rng = np.random.default_rng(100 + seed)
        pts = (rng.integers(0, 4000, (600, 2)).astype(np.float32) * np.float32(0.37))
        for thr in (20.0, 55.5):
            ref = oracle.cluster_euclidean(pts, thr, 5)
            got = ctx.cluster_points(pts, thr, 5)
            assert got[1] == ref[1] and np.array_equal(got[0], ref[0])
            for a_, b_ in zip(got[2:], ref[2:]):
                assert np.array_equal(a_, b_)


# ---- FlowClusterer::getClusters (flow_clusterer.cpp:178-227): the node's path when egomotion is off (node.cpp:375) ----------
@pytest.mark.parametrize("n,dthr,athr,layout", [(600, 50.0, 0.15, "blobs"), (1500, 90.0, 0.26, "uniform"), (300, 20.0, 3.2, "uniform"),
                                                (1, 50.0, 0.15, "uniform"), (2500, 30.0, 0.15, "field")])
def test_cluster_vectors_identical_to_oracle(capi, oracle, n, dthr, athr, layout):
    rng = np.random.default_rng(n + int(dthr))
    if layout == "blobs":
        c = rng.uniform(50, 600, (5, 2))
        k = rng.integers(0, 5, n)
        pos = c[k] + rng.normal(0, 25, (n, 2))
        dirs = rng.uniform(0, 2 * np.pi, 5)[k] + rng.normal(0, 0.1, n)
        mag = rng.uniform(0.5, 4.0, n)
        vec = np.c_[pos, mag * np.cos(dirs), mag * np.sin(dirs)]
    elif layout == "field":
        # what the node hands over: a Vec4d field on the pixel_step grid, zero vectors skipped, row-major traversal
        h, w, ps = 480, 640, 10
        flow = np.zeros((h, w, 4))
        yy, xx = np.mgrid[0:h:ps, 0:w:ps]
        flow[::ps, ::ps, 0] = xx; flow[::ps, ::ps, 1] = yy
        moving = ((xx - 200) ** 2 + (yy - 150) ** 2 < 90 ** 2) | ((xx - 450) ** 2 + (yy - 300) ** 2 < 70 ** 2)
        flow[::ps, ::ps, 2] = np.where(moving, 2.0 + 0.3 * rng.standard_normal(xx.shape), 0.0)
        flow[::ps, ::ps, 3] = np.where(moving, -1.0 + 0.3 * rng.standard_normal(xx.shape), 0.0)
        vec = oracle.flow_field_vectors(flow, ps)
        assert 50 < len(vec) < 1000
    else:
        vec = np.c_[rng.uniform(0, 640, (n, 2)), rng.normal(0, 2, (n, 2))]
    ctx = _ctx(capi, 640, 480)
    ref_lab, ref_n = oracle.cluster_vectors(vec, dthr, athr)
    lab, ncl = ctx.cluster_vectors(vec, dthr, athr)
    assert ncl == ref_n
    assert np.array_equal(lab, ref_lab)
    # deterministic
    lab2, _ = ctx.cluster_vectors(vec, dthr, athr)
    assert np.array_equal(lab, lab2)


def test_cluster_vectors_empty(capi):
    ctx = _ctx(capi, 640, 480)
    lab, ncl = ctx.cluster_vectors(np.zeros((0, 4)))
    assert len(lab) == 0 and ncl == 0
