"""The oracle's restatement against the reference's OWN code.

oracle/_ref holds the reference's unmodified common/src/VarFlow.cpp and common/src/{flow_clusterer,vector_cluster,
point_cluster}.cpp, compiled where they lie under /root/reference against small header shims (oracle/ref_shim: the legacy
OpenCV C API on the oracle's cv2-pinned primitives; a minimal cv::Mat / Point2f / Vec4d).  The in-tree arithmetic of SURVEY
rows a11-a14 (Gauss-Seidel step / iteration / recursion, residual, CalcFlow schedule) and of 8f-3 (clusterEuclidean,
getClusters) is thereby pinned to the reference's loops, bit for bit.

/root/reference does not exist on the GPU box: the live comparisons run where oracle/_ref was built (this container; the
library also travels with the snapshot), the frozen ones (tests/golden/golden_ref.npz, written by
tests/golden/make_golden_ref.py from the reference code) run everywhere."""
import os

import numpy as np
import pytest

from motion_detection_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_ref.npz")


def _need(oracle, name):
    try:
        lib = oracle.ref_lib(name)
    except Exception as e:                                   # pragma: no cover
        pytest.skip("oracle/_ref not buildable here: %s" % e)
    if lib is None:
        pytest.skip("oracle/_ref/lib%s_ref.so not built (no /root/reference)" % name)


def _vf_pair(w, h, seed, patch=False):
    if patch:
        return synth.sequence(w, h, 2, seed=seed, camera=False, blobs=0, patch=True)[0]
    return synth.sequence(w, h, 2, seed=seed, camera=False, blobs=0, whole_field=(0.75, -0.5), margin=16)[0]


@pytest.mark.parametrize("size,seed,patch", [((96, 80), 5, False), ((333, 211), 6, False), ((640, 480), 1234, True), ((50, 37), 2, False)])
def test_varflow_restatement_equals_reference_code(oracle, size, seed, patch):
    """VarFlow::CalcFlow (VarFlow.cpp:600-697) with the varFlow() parameters (optical_flow_calculator.cpp:422-429)."""
    _need(oracle, "varflow")
    fr = _vf_pair(size[0], size[1], seed, patch)
    U, V = oracle.varflow(fr[0], fr[1])
    Ur, Vr = oracle.ref_varflow(fr[0], fr[1])
    assert np.array_equal(U, Ur) and np.array_equal(V, Vr), (np.abs(U - Ur).max(), np.abs(V - Vr).max())
    assert np.abs(Ur).max() > 0.05


@pytest.mark.parametrize("params", [dict(max_level=2, n1=1, n2=3), dict(max_level=6, n1=2, n2=2, alpha=700.0), dict(max_level=0),
                                    dict(rho=1.2, sigma=0.8)])
def test_varflow_other_parameters_equal_reference_code(oracle, params):
    """other multigrid depths / sweep counts / smoothing: the schedule of gauss_seidel_recursive (VarFlow.cpp:508-584), the
    max_level clamp of the constructor (:68-83) and the kernel sizes of cvSmooth follow the reference for any parameters"""
    _need(oracle, "varflow")
    fr = _vf_pair(160, 120, 9)
    U, V = oracle.varflow(fr[0], fr[1], **params)
    Ur, Vr = oracle.ref_varflow(fr[0], fr[1], **params)
    assert np.array_equal(U, Ur) and np.array_equal(V, Vr)


def _lattice_points(seed, n=700):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 4000, (n, 2)).astype(np.float32) * np.float32(0.37)       # many (near-)equal distances


def test_cluster_euclidean_restatement_equals_reference_code(oracle):
    """FlowClusterer::clusterEuclidean (flow_clusterer.cpp:231-269) + PointCluster (point_cluster.cpp:26-38,62-65), including the
    overload the unqualified sqrt(float) resolves to with this toolchain."""
    _need(oracle, "cluster")
    for seed in range(12):
        pts = _lattice_points(seed)
        for thr in (20.0, 33.3, 55.5):
            lab, ncl, _, _, _ = oracle.cluster_euclidean(pts, thr, 5)
            sizes, mem = oracle.clusters_from_labels(pts, lab, ncl)
            rs, rm = oracle.ref_cluster_euclidean(pts, thr)
            assert np.array_equal(sizes, rs) and np.array_equal(mem, rm), (seed, thr)


def _flow_field(seed, w=320, h=240, ps=10, frac=0.5):
    rng = np.random.default_rng(seed)
    flow = np.zeros((h, w, 4))
    yy, xx = np.mgrid[0:h:ps, 0:w:ps]
    flow[::ps, ::ps, 0] = xx
    flow[::ps, ::ps, 1] = yy
    mv = rng.random(xx.shape) < frac
    flow[::ps, ::ps, 2] = np.where(mv, rng.normal(0, 2, xx.shape), 0)
    flow[::ps, ::ps, 3] = np.where(mv, rng.normal(0, 2, xx.shape), 0)
    return flow


def test_get_clusters_restatement_equals_reference_code(oracle):
    """FlowClusterer::getClusters (flow_clusterer.cpp:178-227) + VectorCluster (vector_cluster.cpp:25-50,116-136), including the
    overload the unqualified abs(double) resolves to with this toolchain."""
    _need(oracle, "cluster")
    for seed in range(6):
        flow = _flow_field(seed)
        for dt, at in ((50.0, 0.15), (30.0, 0.5), (90.0, 0.26), (25.0, 2.0)):
            vec = oracle.flow_field_vectors(flow, 10)
            lab, ncl = oracle.cluster_vectors(vec, dt, at)
            sizes, mem = oracle.clusters_from_labels(vec, lab, ncl)
            rs, rm = oracle.ref_get_clusters(flow, 10, dt, at)
            assert np.array_equal(sizes, rs) and np.array_equal(mem, rm), (seed, dt, at)
            assert len(rs) > 0 or at < 0.2


# ---- frozen outputs of the reference code (run everywhere, including the GPU box) ------------------------------------------------
def test_oracle_equals_frozen_reference_outputs(oracle):
    G = np.load(GOLD)
    fr = _vf_pair(96, 80, 5)
    U, V = oracle.varflow(fr[0], fr[1])
    assert np.array_equal(U, G["vf_U_96x80"]) and np.array_equal(V, G["vf_V_96x80"])
    fr = _vf_pair(160, 120, 9)
    U, V = oracle.varflow(fr[0], fr[1], max_level=2, n1=1, n2=3)
    assert np.array_equal(U, G["vf_U_160x120_l2"]) and np.array_equal(V, G["vf_V_160x120_l2"])
    pts = _lattice_points(3)
    lab, ncl, _, _, _ = oracle.cluster_euclidean(pts, 33.3, 5)
    sizes, mem = oracle.clusters_from_labels(pts, lab, ncl)
    assert np.array_equal(sizes, G["ce_sizes"]) and np.array_equal(mem, G["ce_members"])
    flow = _flow_field(2)
    vec = oracle.flow_field_vectors(flow, 10)
    lab, ncl = oracle.cluster_vectors(vec, 50.0, 0.5)
    sizes, mem = oracle.clusters_from_labels(vec, lab, ncl)
    assert np.array_equal(sizes, G["gc_sizes"]) and np.array_equal(mem, G["gc_members"])


# ---- OpticalFlowCalculator: the reference's own composition (oracle/_ref/libofc_ref.so) ----------------------------------------------
def _texture(w, h, seed):
    """high-contrast smooth texture (trackable everywhere)"""
    import cv2
    rng = np.random.default_rng(seed)
    t = cv2.GaussianBlur(rng.random((h, w)).astype(np.float32), (0, 0), 2.0)
    t = (t - t.min()) / (t.max() - t.min())
    return np.clip((t - 0.5) * 6.0 + 0.5, 0.0, 1.0)


def _mover_pair(w=512, h=384, ps=64, seed=11):
    """A static textured scene in which a few 70 x 70 patches, each centred on ONE grid point of a 64-pixel grid, move by a few
    pixels: the vectors that survive the min_vector_size filter are those grid points, in different grid columns and not
    collinear -- so the reference's literal getPerspectiveTransform of the first four kept vectors (cpp:120) is well posed."""
    base = _texture(w, h, seed)
    f0 = (base * 255).astype(np.uint8)
    f1 = f0.copy()
    movers = [((64, 64), (3, 2)), ((128, 256), (-2, 3)), ((256, 128), (4, -2)), ((320, 320), (2, 2)), ((448, 192), (-3, -3))]
    for (cx, cy), (dx, dy) in movers:
        x0, y0 = cx - 35, cy - 35
        f1[y0 + dy:y0 + dy + 70, x0 + dx:x0 + dx + 70] = f0[y0:y0 + 70, x0:x0 + 70]
    return f0, f1, ps


def test_chain_composition_equals_reference_code(oracle):
    """OpticalFlowCalculator::calculateOpticalFlow (optical_flow_calculator.cpp:30-130), unmodified: grid order, status /
    min_vector_size filter, the Vec4d field, getPerspectiveTransform of the first four kept vectors, warpPerspective, absdiff,
    threshold 190 -- against the oracle's composition in its literal mode (MODE_FIRST4, no morphology)."""
    _need(oracle, "ofc")
    f0, f1, ps = _mover_pair()
    h, w = f0.shape
    for mvs in (1.0, 0.2):
        nv, flow, comp = oracle.ref_calculate_optical_flow(f0, f1, ps, mvs)
        o = oracle.process_pair(f0, f1, pixel_step=ps, min_vector_size=mvs, mode=oracle.MODE_FIRST4, morph=False)
        assert nv == o["num_vectors"] and nv >= 4
        assert np.array_equal(flow, oracle.flow_field_from_filter(o["pts"], o["status"], o["flow4"], w, h))
        assert o["inliers"] == 4, "the first four kept vectors must be in general position for this test"
        assert comp is not None and np.array_equal(comp, o["mask"])
        if mvs == 1.0:
            assert 0 < int((comp > 0).sum()) < comp.size // 4      # the comparison is not vacuous


def test_chain_without_vectors_leaves_comp_untouched_like_the_reference(oracle):
    """num_vectors == 0 (static scene): the reference skips the warp and leaves comp as it was (cpp:118); the field holds the grid
    points with zero vectors"""
    _need(oracle, "ofc")
    f0 = (_texture(256, 192, 3) * 255).astype(np.uint8)
    nv, flow, comp = oracle.ref_calculate_optical_flow(f0, f0, 32, 1.0)
    o = oracle.process_pair(f0, f0, pixel_step=32, min_vector_size=1.0, mode=oracle.MODE_FIRST4, morph=False)
    assert nv == 0 == o["num_vectors"] and comp is None and int(o["mask"].sum()) == 0
    assert np.array_equal(flow, oracle.flow_field_from_filter(o["pts"], o["status"], o["flow4"], 256, 192))


def test_trajectory_bookkeeping_equals_reference_code(oracle):
    """calculateOpticalFlowTrajectory (cpp:133-257), unmodified: per-pair LK from the tracked positions, the 10-pixel border rule,
    complete trajectories in grid order, the Vec4d field of the LAST pair."""
    _need(oracle, "ofc")
    frames, _ = synth.sequence(320, 240, 5, seed=21)
    for ps, mvs in ((20, 1.0), (16, 0.2)):
        nv, flow, traj = oracle.ref_calculate_trajectories(frames, ps, mvs)
        tr, ln, last = oracle.track_trajectories(frames, ps)
        full = tr[ln == len(frames)]
        assert traj.shape == full.shape and len(full) > 50 and np.array_equal(traj, full)
        prev, nxt, st = last
        nv_o, keep, flow4 = oracle.flow_filter(prev, nxt, st, mvs)
        assert nv == nv_o
        # the field is indexed by the START point of the last pair (cpp:187,215): tracked positions, truncated to int
        exp = np.zeros_like(flow)
        for p, row in zip(prev, flow4):
            exp[int(p[1]), int(p[0])] = row
        assert np.array_equal(flow, exp)


def test_compensated_flow_equals_reference_code(oracle):
    """calculateCompensatedFlow (cpp:264-335), unmodified: image-input LK with MAX_LEVEL 2, fixed 1.0 filter"""
    _need(oracle, "ofc")
    frames, _ = synth.sequence(320, 240, 2, seed=5)
    flow = oracle.ref_calculate_compensated_flow(frames[0], frames[1], 10)
    pts = oracle.grid_points(320, 240, 10)
    p2, st = oracle.lk(frames[0], frames[1], pts, max_level=2)
    _, _, flow4 = oracle.flow_filter(pts, p2, st, 1.0)
    assert np.array_equal(flow, oracle.flow_field_from_filter(pts, st, flow4, 320, 240))
    assert (np.abs(flow[..., 2:]) > 0).any()


def test_write_flow_and_trajectories_formats_are_the_reference_code(oracle, tmp_path):
    """writeFlow / writeTrajectories (cpp:509-562), unmodified, on the very inputs adapter/test/node_callsites.cpp feeds the adapter's
    host implementations (tests/test_adapter_cpu.py expects the same lines from those)."""
    _need(oracle, "ofc")
    flow = np.zeros((20, 30, 4))
    for y in range(0, 20, 10):
        for x in range(0, 30, 10):
            flow[y, x, 0], flow[y, x, 1] = x, y
    flow[0, 10, 2:] = (1.5, -0.25)
    flow[10, 20, 2:] = (-2.0, 0.333333333)
    flow[10, 0] = (-1.0, -1.0, 7.0, 7.0)
    oracle.ref_write_flow(flow, 10, str(tmp_path / "flow"))
    assert open(tmp_path / "flow_h").read().splitlines() == ["0, 1.5, 0", "0, 0, -2"]
    assert open(tmp_path / "flow_f").read().splitlines() == ["0, -0.25, 0", "0, 0, 0.333333"]
    traj = np.array([[[10, 20], [11.25, 19.5], [12.5, 19]]], np.float32)
    oracle.ref_write_trajectories(traj, str(tmp_path / "traj"))
    assert open(tmp_path / "traj").read().splitlines() == ["10, 20, 11.25, 19.5, 12.5, 19"]


def test_frozen_reference_chain_outputs(oracle):
    """the same two comparisons against outputs of the reference code frozen in tests/golden/golden_ref.npz (runs where
    /root/reference does not exist)"""
    G = np.load(GOLD)
    f0, f1, ps = _mover_pair()
    o = oracle.process_pair(f0, f1, pixel_step=ps, min_vector_size=1.0, mode=oracle.MODE_FIRST4, morph=False)
    h, w = f0.shape
    assert o["num_vectors"] == int(G["ofc_nv"][0])
    assert np.array_equal(oracle.flow_field_from_filter(o["pts"], o["status"], o["flow4"], w, h)[::ps, ::ps], G["ofc_flow_grid"])
    assert np.array_equal(np.packbits(o["mask"] > 0), G["ofc_comp_bits"])
    frames, _ = synth.sequence(320, 240, 5, seed=21)
    tr, ln, last = oracle.track_trajectories(frames, 20)
    assert np.array_equal(tr[ln == len(frames)], G["traj_complete"])
    assert oracle.flow_filter(last[0], last[1], last[2], 1.0)[0] == int(G["traj_nv"][0])


# ---- OutlierDetector: the reference's own code (oracle/_ref/libod_ref.so) ----------------------------------------------------------
def _mad_field(seed, w=320, h=240, ps=10, zero_frac=0.3):
    rng = np.random.default_rng(seed)
    flow = np.zeros((h, w, 4))
    for y in range(0, h, ps):
        for x in range(0, w, ps):
            if rng.random() < zero_frac:
                flow[y, x] = (x, y, 0.0, 0.0)
            else:
                d = rng.normal((2.0, -1.0), 0.3) if rng.random() > 0.08 else rng.normal((-4.0, 6.0), 1.5)
                flow[y, x] = (x, y, d[0], d[1])
    return flow


@pytest.mark.parametrize("include_zeros", [False, True])
def test_find_outliers_restatement_equals_reference_code(oracle, include_zeros):
    """OutlierDetector::findOutliers / createMask / getMedian (outlier_detector.cpp:37-186), unmodified: the angle stage and the
    magnitude stage share one mask; even-count medians average the two middle values"""
    _need(oracle, "od")
    for seed in range(6):
        flow = _mad_field(seed, ps=10, zero_frac=0.3 if seed % 2 else 0.0)
        prob = oracle.ref_find_outliers(flow, 10, include_zeros)
        nodes = flow[::10, ::10].reshape(-1, 4)
        out, _ = oracle.find_outliers(nodes[:, 2:4], include_zeros)
        assert np.array_equal(prob[::10, ::10].reshape(-1) == 1.0, out == 1), (seed, include_zeros)
        assert 0 < int(out.sum()) < len(out)


def _two_motion_trajectories(seed, T=300, F=5, n_out=30):
    """points of a small region (centroid-relative coordinates of a few tens of pixels: the f32 rounding of the reference's residuals
    stays far below the threshold) carried by one affine camera motion plus 0.01 px noise; n_out of them jitter by +-3 px per frame"""
    rng = np.random.default_rng(seed)
    p0 = np.stack(np.meshgrid(np.arange(100, 160, 3.0), np.arange(80, 125, 3.0)), -1).reshape(-1, 2)[:T]
    traj = np.zeros((len(p0), F, 2), np.float32)
    out = rng.choice(len(p0), n_out, replace=False)
    A = np.array([[1.002, 0.001], [-0.001, 0.999]]); t = np.array([1.5, -0.7])
    cur = p0.copy()
    for j in range(F):
        traj[:, j] = cur + rng.normal(0, 0.01, cur.shape)
        traj[out, j] += rng.choice([-3.0, 3.0], (n_out, 2))
        cur = cur @ A.T + t
    return traj, np.sort(out)


def test_fit_subspace_composition_follows_reference_code(oracle):
    """OutlierDetector::fitSubspace (outlier_detector.cpp:236-331), unmodified, with a small Eigen stand-in under it (f32 matrices in
    index order, Jacobi SVD): data layout, mean subtraction, rand() % T sampling after srand(seed), projector, residual, inlier count,
    first-best rule, chi-square threshold, and which point of a trajectory is reported -- on well separated data the winning sample
    and the reported outliers are those of the oracle (whose sums are f64: not a bit-level comparison)"""
    _need(oracle, "od")
    for seed in (1, 7, 23):
        traj, planted = _two_motion_trajectories(seed)
        oi, cols = oracle.ref_fit_subspace(traj, 2, 0.5, seed)
        n, res, ocols, outl, thr = oracle.fit_subspace(traj, 2, 0.5, seed)
        assert np.array_equal(cols, ocols), seed
        assert np.array_equal(oi, np.nonzero(outl)[0]), seed
        # a planted trajectory that was drawn into the winning sample spans the subspace itself: residual 0, not reported
        assert set(oi.tolist()) <= set(planted.tolist()) and len(oi) >= len(planted) - 8, seed


def test_frozen_reference_outlier_detector_outputs(oracle):
    """findOutliers flags and the fitSubspace winner / outliers of the reference code, frozen in tests/golden/golden_ref.npz"""
    G = np.load(GOLD)
    nodes = _mad_field(3, ps=10, zero_frac=0.3)[::10, ::10].reshape(-1, 4)
    assert np.array_equal(oracle.find_outliers(nodes[:, 2:4], False)[0], G["mad_flags_nozero"])
    assert np.array_equal(oracle.find_outliers(nodes[:, 2:4], True)[0], G["mad_flags_zero"])
    traj, _ = _two_motion_trajectories(7)
    n, res, cols, outl, thr = oracle.fit_subspace(traj, 2, 0.5, 7)
    assert np.array_equal(cols, G["sub_cols"]) and np.array_equal(np.nonzero(outl)[0], G["sub_outliers"])
