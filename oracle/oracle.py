"""ctypes binding of the CPU oracle (libmd_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under motion_detection_b200/ imports this module.
Each function names the reference code it restates (see md_oracle*.c headers for file:line).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u8p = C.POINTER(C.c_uint8)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)
i32p = C.POINTER(C.c_int32)
i16p = C.POINTER(C.c_int16)


def build(force=False):
    """gcc-compile the restatement (oracle/Makefile)."""
    so = os.path.join(_HERE, "libmd_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libmd_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_glibc_rand.restype = C.c_int
        _LIB.orc_pyr_levels.restype = C.c_int
        _LIB.orc_lk_pyr.restype = C.c_int
        _LIB.orc_grid_points.restype = C.c_int
        _LIB.orc_flow_filter.restype = C.c_int
        _LIB.orc_perspective_4pt.restype = C.c_int
        _LIB.orc_fit_egomotion.restype = C.c_int
        _LIB.orc_varflow.restype = C.c_int
        _LIB.orc_fit_subspace.restype = C.c_int
        _LIB.orc_cluster_euclidean.restype = C.c_int
        _LIB.orc_bounding_boxes.restype = C.c_int
        _LIB.orc_find_outliers.restype = C.c_int
        _LIB.orc_cluster_vectors.restype = C.c_int
    return _LIB


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(u8p)


def glibc_rand(seed, n):
    """srand(seed); n x rand()  (outlier_detector.cpp:17,228)."""
    st = (C.c_int32 * 36)()
    lib().orc_glibc_srand(st, C.c_uint32(seed))
    return [lib().orc_glibc_rand(st) for _ in range(n)]


def gray(rgb):
    rgb, p = _u8(rgb)
    h, w, _ = rgb.shape
    out = np.empty((h, w), np.uint8)
    lib().orc_gray_bgr2gray(p, w, h, w * 3, out.ctypes.data_as(u8p), w)
    return out


def pyr_down(img):
    img, p = _u8(img)
    h, w = img.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().orc_pyr_down(p, w, h, w, out.ctypes.data_as(u8p), out.shape[1])
    return out


def pyr_levels(w, h, win=40, max_level=5):
    return lib().orc_pyr_levels(w, h, win, max_level)


def pyramid(img, win=40, max_level=5):
    lv = [np.ascontiguousarray(img, np.uint8)]
    for _ in range(pyr_levels(img.shape[1], img.shape[0], win, max_level)):
        lv.append(pyr_down(lv[-1]))
    return lv


def scharr(img):
    img, p = _u8(img)
    h, w = img.shape
    out = np.empty((h, w, 2), np.int16)
    lib().orc_scharr(p, w, h, w, out.ctypes.data_as(i16p))
    return out


def grid_points(w, h, ps):
    n = lib().orc_grid_points(w, h, ps, None)
    pts = np.empty((n, 2), np.float32)
    lib().orc_grid_points(w, h, ps, pts.ctypes.data_as(f32p))
    return pts


def lk(prev, nxt, pts, win=40, max_level=5, max_iters=10, eps=0.03, min_eig=1e-3):
    """cv::calcOpticalFlowPyrLK as called at optical_flow_calculator.cpp:71."""
    prev, pp = _u8(prev)
    nxt, np_ = _u8(nxt)
    h, w = prev.shape
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    out = np.zeros_like(pts)
    st = np.zeros(len(pts), np.uint8)
    lib().orc_lk_pyr(pp, np_, w, h, w, pts.ctypes.data_as(f32p), len(pts), out.ctypes.data_as(f32p),
                     st.ctypes.data_as(u8p), win, max_level, max_iters, C.c_double(eps), C.c_float(min_eig))
    return out, st


def flow_filter(p1, p2, status, min_vec):
    p1 = np.ascontiguousarray(p1, np.float32)
    p2 = np.ascontiguousarray(p2, np.float32)
    status = np.ascontiguousarray(status, np.uint8)
    n = len(status)
    keep = np.zeros(n, np.uint8)
    flow4 = np.zeros((n, 4), np.float64)
    nv = lib().orc_flow_filter(p1.ctypes.data_as(f32p), p2.ctypes.data_as(f32p), status.ctypes.data_as(u8p), n,
                               C.c_double(min_vec), keep.ctypes.data_as(u8p), flow4.ctypes.data_as(f64p))
    return nv, keep, flow4


def perspective_4pt(src, dst):
    src = np.ascontiguousarray(src, np.float64)
    dst = np.ascontiguousarray(dst, np.float64)
    H = np.zeros(9, np.float64)
    ok = lib().orc_perspective_4pt(src.ctypes.data_as(f64p), dst.ctypes.data_as(f64p), H.ctypes.data_as(f64p))
    return ok, H.reshape(3, 3)


MODE_FIRST4, MODE_RANSAC_HOMOGRAPHY, MODE_RANSAC_AFFINE = 0, 1, 2


def fit_egomotion(p1, p2, keep, w, h, mode=MODE_RANSAC_HOMOGRAPHY, iters=50, thr=0.5, seed=1):
    p1 = np.ascontiguousarray(p1, np.float32)
    p2 = np.ascontiguousarray(p2, np.float32)
    keep = np.ascontiguousarray(keep, np.uint8)
    H = np.zeros(9, np.float64)
    inl = np.zeros(len(keep), np.uint8)
    n = lib().orc_fit_egomotion(p1.ctypes.data_as(f32p), p2.ctypes.data_as(f32p), keep.ctypes.data_as(u8p), len(keep),
                                w, h, mode, iters, C.c_double(thr), C.c_uint32(seed), H.ctypes.data_as(f64p),
                                inl.ctypes.data_as(u8p))
    return n, H.reshape(3, 3), inl


def warp_perspective(src, H):
    src, p = _u8(src)
    h, w = src.shape
    H = np.ascontiguousarray(H, np.float64)
    out = np.empty_like(src)
    lib().orc_warp_perspective(p, w, h, w, H.ctypes.data_as(f64p), out.ctypes.data_as(u8p), w)
    return out


def absdiff_threshold(a, b, thresh=190):
    a, pa = _u8(a)
    b, pb = _u8(b)
    h, w = a.shape
    out = np.empty_like(a)
    lib().orc_absdiff_threshold(pa, pb, w, h, w, thresh, out.ctypes.data_as(u8p), w)
    return out


def erode3(a):
    a, pa = _u8(a)
    h, w = a.shape
    out = np.empty_like(a)
    lib().orc_erode3(pa, w, h, w, out.ctypes.data_as(u8p), w)
    return out


def dilate3(a):
    a, pa = _u8(a)
    h, w = a.shape
    out = np.empty_like(a)
    lib().orc_dilate3(pa, w, h, w, out.ctypes.data_as(u8p), w)
    return out


def motion_mask(prev, cur, H, thresh=190, morph=True):
    prev, pp = _u8(prev)
    cur, pc = _u8(cur)
    h, w = prev.shape
    H = np.ascontiguousarray(H, np.float64)
    out = np.empty_like(prev)
    lib().orc_motion_mask(pp, pc, w, h, w, H.ctypes.data_as(f64p), thresh, 1 if morph else 0, out.ctypes.data_as(u8p), w)
    return out


def process_pair(prev, cur, pixel_step=10, min_vector_size=0.2, mode=MODE_RANSAC_HOMOGRAPHY, iters=50, thr=0.5,
                 seed=1, thresh=190, morph=True, lk_kwargs=None):
    """The oracle composition: grid LK -> filter -> H -> warp -> absdiff -> threshold -> erode -> dilate.
    Follows OpticalFlowCalculator::calculateOpticalFlow (optical_flow_calculator.cpp:30-130) +
    BackgroundSubtractor morphology (background_subtractor.cpp:31-32)."""
    h, w = prev.shape
    pts = grid_points(w, h, pixel_step)
    p2, st = lk(prev, cur, pts, **(lk_kwargs or {}))
    nv, keep, flow4 = flow_filter(pts, p2, st, min_vector_size)
    ninl, H, inl = fit_egomotion(pts, p2, keep, w, h, mode, iters, thr, seed)
    if ninl > 0:
        mask = motion_mask(prev, cur, H, thresh, morph)
    else:
        mask = np.zeros_like(prev)
    return dict(pts=pts, next=p2, status=st, keep=keep, flow4=flow4, num_vectors=nv, H=H, inliers=ninl,
                inlier_mask=inl, mask=mask)


def process_pair_varflow(prev, cur, pixel_step=10, min_vector_size=0.2, mode=MODE_RANSAC_HOMOGRAPHY, iters=50, thr=0.5,
                         seed=1, thresh=190, morph=True):
    """The same composition with the dense variational flow (VarFlow::CalcFlow, VarFlow.cpp:600-697) as the flow engine:
    the field is sampled at the grid points, next = (x + U, y - V) since V is y-up (VarFlow.cpp:103-107)."""
    h, w = prev.shape
    pts = grid_points(w, h, pixel_step)
    U, V = varflow(prev, cur)
    xi = pts[:, 0].astype(np.int64)
    yi = pts[:, 1].astype(np.int64)
    p2 = np.stack([pts[:, 0] + U[yi, xi], pts[:, 1] - V[yi, xi]], axis=1).astype(np.float32)
    st = np.ones(len(pts), np.uint8)
    nv, keep, flow4 = flow_filter(pts, p2, st, min_vector_size)
    ninl, H, inl = fit_egomotion(pts, p2, keep, w, h, mode, iters, thr, seed)
    mask = motion_mask(prev, cur, H, thresh, morph) if ninl > 0 else np.zeros_like(prev)
    return dict(pts=pts, next=p2, status=st, keep=keep, flow4=flow4, num_vectors=nv, H=H, inliers=ninl, inlier_mask=inl,
                mask=mask, U=U, V=V)


def gaussian_blur_f32(img, sigma):
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    out = np.empty_like(img)
    lib().orc_gaussian_blur_f32(img.ctypes.data_as(f32p), w, h, out.ctypes.data_as(f32p), C.c_double(sigma))
    return out


def resize_linear_f32(img, dw, dh):
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    out = np.empty((dh, dw), np.float32)
    lib().orc_resize_linear_f32(img.ctypes.data_as(f32p), w, h, out.ctypes.data_as(f32p), dw, dh)
    return out


def varflow(A, B, max_level=4, start_level=0, n1=2, n2=2, rho=2.8, alpha=1400.0, sigma=1.5, literal=True):
    """VarFlow::CalcFlow with the parameters of OpticalFlowCalculator::varFlow (optical_flow_calculator.cpp:422-429)."""
    A, pa = _u8(A)
    B, pb = _u8(B)
    h, w = A.shape
    U = np.zeros((h, w), np.float32)
    V = np.zeros((h, w), np.float32)
    r = lib().orc_varflow(pa, pb, w, h, w, max_level, start_level, n1, n2, C.c_float(rho), C.c_float(alpha),
                          C.c_float(sigma), U.ctypes.data_as(f32p), V.ctypes.data_as(f32p), 1 if literal else 0)
    if r != 1:
        raise RuntimeError("orc_varflow failed")
    return U, V


def fit_subspace(traj, num_motions=2, sigma=0.5, seed=1, forced_cols=None, iters=50):
    """OutlierDetector::fitSubspace (outlier_detector.cpp:236-331). traj: [T][F][2] f32."""
    traj = np.ascontiguousarray(traj, np.float32)
    T, F, _ = traj.shape
    d = 4 * num_motions
    res = np.zeros(T, np.float32)
    cols = np.zeros(d, np.int32)
    outl = np.zeros(T, np.uint8)
    thr = C.c_double(0)
    fc = None
    if forced_cols is not None:
        fc_arr = np.ascontiguousarray(forced_cols, np.int32).reshape(-1, d)
        iters = fc_arr.shape[0]
        fc = fc_arr.ctypes.data_as(i32p)
    n = lib().orc_fit_subspace(traj.ctypes.data_as(f32p), T, F, num_motions, C.c_double(sigma), C.c_uint32(seed), fc,
                               iters, res.ctypes.data_as(f32p), cols.ctypes.data_as(i32p), outl.ctypes.data_as(u8p),
                               C.byref(thr))
    return n, res, cols, outl, thr.value


# ---- the node's live path (imageCallback, ros/src/motion_detection_node.cpp:235-414) ------------------------------
def track_trajectories(frames, pixel_step=10):
    """calculateOpticalFlowTrajectory (cpp:133-257) on gray frames: returns (traj [P][F][2], len [P], last_prev, last_next,
    last_status)."""
    frames = [np.ascontiguousarray(f, np.uint8) for f in frames]
    F = len(frames)
    h, w = frames[0].shape
    pts = grid_points(w, h, pixel_step)
    P = len(pts)
    cur = pts.copy()
    traj = np.zeros((P, F, 2), np.float32)
    traj[:, 0] = pts
    ln = np.ones(P, np.int32)
    last = None
    for j in range(F - 1):
        nxt, st = lk(frames[j], frames[j + 1], cur)
        if j == F - 2:
            last = (cur.copy(), nxt.copy(), st.copy())
        nxt = np.ascontiguousarray(nxt, np.float32)
        st = np.ascontiguousarray(st, np.uint8)
        lib().orc_traj_step(cur.ctypes.data_as(f32p), nxt.ctypes.data_as(f32p), st.ctypes.data_as(u8p), traj.ctypes.data_as(f32p),
                            ln.ctypes.data_as(i32p), P, F, w, h)
    return traj, ln, last


def cluster_euclidean(pts, distance_threshold=50.0, min_size=5):
    """FlowClusterer::clusterEuclidean + showBoundingBoxes: (labels, nclusters_all, boxes [K][4], sizes [K], ids [K])."""
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    n = len(pts)
    labels = np.zeros(max(n, 1), np.int32)
    ncl = lib().orc_cluster_euclidean(pts.ctypes.data_as(f32p), n, C.c_double(distance_threshold), labels.ctypes.data_as(i32p))
    boxes = np.zeros((max(ncl, 1), 4), np.int32)
    sizes = np.zeros(max(ncl, 1), np.int32)
    ids = np.zeros(max(ncl, 1), np.int32)
    k = lib().orc_bounding_boxes(pts.ctypes.data_as(f32p), n, labels.ctypes.data_as(i32p), ncl, min_size, boxes.ctypes.data_as(i32p),
                                 sizes.ctypes.data_as(i32p), ids.ctypes.data_as(i32p))
    return labels[:n], ncl, boxes[:k], sizes[:k], ids[:k]


def cluster_vectors(vec4, distance_threshold=50.0, angular_threshold=0.15):
    """FlowClusterer::getClusters (flow_clusterer.cpp:178-227) on the participating vectors [n][4] = (x, y, dx, dy) in the
    reference's traversal order: (labels [n], number of clusters founded)."""
    vec4 = np.ascontiguousarray(vec4, np.float64).reshape(-1, 4)
    n = len(vec4)
    labels = np.zeros(max(n, 1), np.int32)
    ncl = lib().orc_cluster_vectors(vec4.ctypes.data_as(f64p), n, C.c_double(distance_threshold), C.c_double(angular_threshold),
                                    labels.ctypes.data_as(i32p))
    return labels[:n], ncl


def flow_field_vectors(flow, pixel_step):
    """The vectors getClusters visits (flow_clusterer.cpp:180-185): rows outer, columns inner, |dx| > 0 or |dy| > 0.
    flow: [h][w][4] f64 (the Vec4d field)."""
    sub = flow[::pixel_step, ::pixel_step].reshape(-1, 4)
    return np.ascontiguousarray(sub[(np.abs(sub[:, 2]) > 0.0) | (np.abs(sub[:, 3]) > 0.0)])


def live_detect(frames, pixel_step=10, num_motions=2, sigma=0.5, distance_threshold=50.0, seed=1, iters=50, min_size=5):
    """The callback body over F = 2 * num_motions + 1 gray frames: trajectories -> fitSubspace -> outlier points ->
    clusterEuclidean -> bounding boxes (node.cpp:294-395)."""
    F = 2 * num_motions + 1
    assert len(frames) == F
    traj, ln, _ = track_trajectories(frames, pixel_step)
    idx = np.nonzero(ln == F)[0].astype(np.int32)
    tc = np.ascontiguousarray(traj[idx])
    out = dict(traj=tc, traj_index=idx, num_trajectories=len(idx))
    if len(idx) == 0:
        return out
    n, res, cols, outl, thr = fit_subspace(tc, num_motions=num_motions, sigma=sigma, seed=seed, iters=iters)
    opts = np.ascontiguousarray(tc[outl != 0][:, F - 2])
    labels, ncl, boxes, sizes, ids = cluster_euclidean(opts, distance_threshold, min_size)
    out.update(subspace_inliers=n, residual=res, best_cols=cols, outlier=outl, outlier_points=opts, labels=labels,
               num_clusters_all=ncl, boxes=boxes, cluster_sizes=sizes, cluster_ids=ids)
    return out


def find_outliers(dxdy, include_zeros=False):
    """OutlierDetector::findOutliers (outlier_detector.cpp:37-186): (outlier u8 [n], stats [median/MAD angle, median/MAD mag])."""
    dxdy = np.ascontiguousarray(dxdy, np.float64).reshape(-1, 2)
    n = len(dxdy)
    out = np.zeros(n, np.uint8)
    stats = np.zeros(4, np.float64)
    lib().orc_find_outliers(dxdy.ctypes.data_as(f64p), n, 1 if include_zeros else 0, out.ctypes.data_as(u8p), stats.ctypes.data_as(f64p))
    return out, stats


# ---- oracle/_ref: the reference's OWN sources compiled against the shims of oracle/ref_shim (recipe: oracle/Makefile) --------
# Used only to validate the restatement above (tests/test_oracle_ref.py) and to generate tests/golden/golden_ref.npz.
_REF = {}


def line_aa(img, p1, p2, colour):
    """cv::line(img, p1, p2, colour, 1, CV_AA, 0) in place; img [h][w] or [h][w][3] uint8."""
    assert img.dtype == np.uint8 and img.flags.c_contiguous
    h, w = img.shape[:2]
    nch = 1 if img.ndim == 2 else img.shape[2]
    col = np.ascontiguousarray(np.array(colour, np.uint8).reshape(-1)[:nch])
    lib().orc_line_aa(img.ctypes.data_as(u8p), w, h, w * nch, nch, int(p1[0]), int(p1[1]), int(p2[0]), int(p2[1]), col.ctypes.data_as(u8p))
    return img


def arrow_segments(elem, pixel_step, min_vector_size):
    seg = np.zeros(12, np.int32)
    e = np.ascontiguousarray(elem, np.float64)
    ok = lib().orc_arrow_segments(e.ctypes.data_as(f64p), int(pixel_step), C.c_double(min_vector_size), seg.ctypes.data_as(i32p))
    return seg.reshape(3, 4) if ok else None


def draw_flow(image, vec4, pixel_step=10, min_vector_size=0.2, colour=(255, 0, 0)):
    """OpticalFlowVisualizer::showOpticalFlowVectors (optical_flow_visualizer.cpp:23-71): vec4 [n][4] = the non-empty elements
    of the flow field in row-major (y outer, x inner) order.  Returns (image with arrows, number of arrows drawn)."""
    image = np.ascontiguousarray(image, np.uint8)
    h, w = image.shape[:2]
    nch = 1 if image.ndim == 2 else image.shape[2]
    vec4 = np.ascontiguousarray(vec4, np.float64).reshape(-1, 4)
    col = np.ascontiguousarray(np.array(colour, np.uint8).reshape(-1)[:nch])
    out = np.empty_like(image)
    n = lib().orc_draw_flow(image.ctypes.data_as(u8p), w, h, nch, w * nch, vec4.ctypes.data_as(f64p), len(vec4), int(pixel_step),
                            C.c_double(min_vector_size), col.ctypes.data_as(u8p), out.ctypes.data_as(u8p), w * nch)
    return out, n


def flow_field_row_major(pts, nxt, status, keep):
    """The flow-field elements (optical_flow_calculator.cpp:78-117) of the grid points, in the order showOpticalFlowVectors meets
    them: row-major over the image (y outer, x inner)."""
    pts = np.asarray(pts, np.float32); nxt = np.asarray(nxt, np.float32)
    d = (nxt - pts).astype(np.float64)
    vec = np.zeros((len(pts), 4), np.float64)
    ok = np.asarray(status) != 0
    kp = ok & (np.asarray(keep) != 0)
    vec[ok, 0] = pts[ok, 0]; vec[ok, 1] = pts[ok, 1]
    vec[~ok, 0] = -1.0; vec[~ok, 1] = -1.0
    vec[kp, 2] = d[kp, 0]; vec[kp, 3] = d[kp, 1]
    order = np.lexsort((pts[:, 0], pts[:, 1]))
    return vec[order]


def ref_lib(name):
    """'varflow' (common/src/VarFlow.cpp), 'cluster' (flow_clusterer.cpp + vector_cluster.cpp + point_cluster.cpp), 'ofc'
    (optical_flow_calculator.cpp) or 'od' (outlier_detector.cpp); None when oracle/_ref was not built (no /root/reference at build time)."""
    if name not in _REF:
        so = os.path.join(_HERE, "_ref", "lib%s_ref.so" % name)
        if not os.path.exists(so) and os.path.exists("/root/reference/common/src/VarFlow.cpp"):
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
        _REF[name] = C.CDLL(so) if os.path.exists(so) else None
    return _REF[name]


def ref_varflow(A, B, max_level=4, start_level=0, n1=2, n2=2, rho=2.8, alpha=1400.0, sigma=1.5):
    """The reference's VarFlow class itself (ctor + CalcFlow, VarFlow.cpp:27-162,600-697)."""
    A, pa = _u8(A)
    B, pb = _u8(B)
    h, w = A.shape
    U = np.zeros((h, w), np.float32)
    V = np.zeros((h, w), np.float32)
    rc = ref_lib("varflow").ref_varflow(pa, pb, w, h, w, max_level, start_level, n1, n2, C.c_float(rho), C.c_float(alpha),
                                        C.c_float(sigma), U.ctypes.data_as(f32p), V.ctypes.data_as(f32p))
    if rc != 1:
        raise RuntimeError("reference VarFlow::CalcFlow returned %d" % rc)
    return U, V


def ref_cluster_euclidean(pts, distance_threshold):
    """The reference's FlowClusterer::clusterEuclidean: (sizes [K], members [sum sizes][2]) of the clusters it returns."""
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    n = len(pts)
    sizes = np.zeros(n + 1, np.int32)
    mem = np.zeros((max(n, 1), 2), np.float32)
    k = ref_lib("cluster").ref_cluster_euclidean(pts.ctypes.data_as(f32p), n, C.c_double(distance_threshold),
                                                 sizes.ctypes.data_as(i32p), mem.ctypes.data_as(f32p))
    return sizes[:k].copy(), mem[:int(sizes[:k].sum())].copy()


def ref_get_clusters(flow, pixel_step, distance_threshold, angular_threshold):
    """The reference's FlowClusterer::getClusters on a Vec4d field [h][w][4] f64: (sizes [K], members [sum sizes][4])."""
    flow = np.ascontiguousarray(flow, np.float64)
    h, w, _ = flow.shape
    n = ((h + pixel_step - 1) // pixel_step) * ((w + pixel_step - 1) // pixel_step)
    sizes = np.zeros(n + 1, np.int32)
    mem = np.zeros((max(n, 1), 4), np.float64)
    k = ref_lib("cluster").ref_get_clusters(flow.ctypes.data_as(f64p), w, h, pixel_step, C.c_double(distance_threshold),
                                            C.c_double(angular_threshold), sizes.ctypes.data_as(i32p), mem.ctypes.data_as(f64p))
    return sizes[:k].copy(), mem[:int(sizes[:k].sum())].copy()


def clusters_from_labels(items, labels, ncl, min_size=5):
    """What the reference returns (clusters with MORE than min_size members, creation order, members in arrival order) from the
    oracle's labels: (sizes [K], members concatenated)."""
    cnt = np.bincount(labels, minlength=ncl) if len(labels) else np.zeros(0, np.int64)
    ids = [i for i in range(ncl) if cnt[i] > min_size]
    mem = np.concatenate([items[labels == i] for i in ids]) if ids else items[:0]
    return cnt[ids].astype(np.int32), mem


# ---- the reference's own OpticalFlowCalculator (oracle/_ref/libofc_ref.so: common/src/optical_flow_calculator.cpp unmodified, the
# cv:: calls it makes forwarded to the primitives above by oracle/ref_shim_ofc) ------------------------------------------------
def _bgr(img):
    img = np.ascontiguousarray(img, np.uint8)
    if img.ndim == 2:
        img = np.repeat(img[..., None], 3, axis=2)
    return np.ascontiguousarray(img)


def ref_calculate_optical_flow(img1, img2, pixel_step=10, min_vector_size=1.0):
    """OpticalFlowCalculator::calculateOpticalFlow (cpp:30-130): (num_vectors, flow [h][w][4] f64, comp [h][w] u8 or None when the
    reference left comp untouched)."""
    a, b = _bgr(img1), _bgr(img2)
    h, w = a.shape[:2]
    flow = np.zeros((h, w, 4), np.float64)
    comp = np.zeros((h, w), np.uint8)
    rows = C.c_int(0)
    nv = ref_lib("ofc").ref_calculate_optical_flow(a.ctypes.data_as(u8p), b.ctypes.data_as(u8p), w, h, pixel_step,
                                                   C.c_double(min_vector_size), flow.ctypes.data_as(f64p), comp.ctypes.data_as(u8p),
                                                   C.byref(rows))
    return nv, flow, (comp if rows.value == h else None)


def ref_calculate_trajectories(images, pixel_step=10, min_vector_size=1.0):
    """OpticalFlowCalculator::calculateOpticalFlowTrajectory (cpp:133-257): (num_vectors, flow [h][w][4], trajectories [T][F][2])."""
    imgs = np.ascontiguousarray(np.stack([_bgr(i) for i in images]))
    n, h, w = imgs.shape[:3]
    P = ((w + pixel_step - 1) // pixel_step) * ((h + pixel_step - 1) // pixel_step)
    flow = np.zeros((h, w, 4), np.float64)
    traj = np.zeros((P, n, 2), np.float32)
    nt = C.c_int(0)
    nv = ref_lib("ofc").ref_calculate_trajectories(imgs.ctypes.data_as(u8p), n, w, h, pixel_step, C.c_double(min_vector_size),
                                                   flow.ctypes.data_as(f64p), traj.ctypes.data_as(f32p), P, C.byref(nt))
    return nv, flow, traj[:nt.value].copy()


def ref_calculate_compensated_flow(img1, img2, pixel_step=10):
    """OpticalFlowCalculator::calculateCompensatedFlow (cpp:264-335): the flow field [h][w][4]."""
    a, b = _bgr(img1), _bgr(img2)
    h, w = a.shape[:2]
    flow = np.zeros((h, w, 4), np.float64)
    ref_lib("ofc").ref_calculate_compensated_flow(a.ctypes.data_as(u8p), b.ctypes.data_as(u8p), w, h, pixel_step, flow.ctypes.data_as(f64p))
    return flow


def ref_write_flow(flow, pixel_step, filename):
    flow = np.ascontiguousarray(flow, np.float64)
    h, w = flow.shape[:2]
    ref_lib("ofc").ref_write_flow(flow.ctypes.data_as(f64p), w, h, pixel_step, filename.encode())


def ref_write_trajectories(traj, filename):
    traj = np.ascontiguousarray(traj, np.float32)
    ref_lib("ofc").ref_write_trajectories(traj.ctypes.data_as(f32p), traj.shape[0], traj.shape[1], filename.encode())


def flow_field_from_filter(pts, status, flow4, w, h):
    """The Vec4d field calculateOpticalFlow leaves in the node's Mat (cpp:78-117), from the oracle's per-point flow_filter rows."""
    f = np.zeros((h, w, 4), np.float64)
    xi = pts[:, 0].astype(np.int64)
    yi = pts[:, 1].astype(np.int64)
    f[yi, xi] = flow4
    return f


# ---- the reference's own OutlierDetector (oracle/_ref/libod_ref.so: common/src/outlier_detector.cpp unmodified; Eigen is a small
# stand-in, see oracle/ref_shim_ofc/Eigen/Dense) -------------------------------------------------------------------------------
def ref_find_outliers(flow, pixel_step, include_zeros=False):
    """OutlierDetector::findOutliers (outlier_detector.cpp:37-186) on a Vec4d field [h][w][4]: the outlier_probabilities matrix."""
    flow = np.ascontiguousarray(flow, np.float64)
    h, w = flow.shape[:2]
    prob = np.zeros((h, w), np.float64)
    ref_lib("od").ref_find_outliers(flow.ctypes.data_as(f64p), w, h, pixel_step, 1 if include_zeros else 0, prob.ctypes.data_as(f64p))
    return prob


def ref_fit_subspace(traj, num_motions=2, sigma=0.5, seed=1):
    """OutlierDetector::fitSubspace (outlier_detector.cpp:236-331) after srand(seed): (indices of the reported outlier trajectories,
    sampled columns of the winning hypothesis)."""
    traj = np.ascontiguousarray(traj, np.float32)
    T, F, _ = traj.shape
    oi = np.zeros(T, np.int32)
    cols = np.zeros(4 * num_motions, np.int32)
    nc = C.c_int(0)
    n = ref_lib("od").ref_fit_subspace(traj.ctypes.data_as(f32p), T, F, num_motions, C.c_double(sigma), C.c_uint32(seed),
                                       oi.ctypes.data_as(i32p), cols.ctypes.data_as(i32p), C.byref(nc))
    if n < 0:
        raise RuntimeError("could not map the reported outlier points back to trajectories (duplicate points)")
    return oi[:n].copy(), cols[:nc.value].copy()
