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
