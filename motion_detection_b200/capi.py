"""ctypes binding of libmotion_b200.so (include/motion_b200.h).

This is the ONLY compute path of the package: if the CUDA library is missing or no sm_100 GPU is usable the
calls raise -- there is no CPU fallback and nothing here imports oracle/.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmotion_b200.so")

MD_OK = 0
MD_MEM_HOST, MD_MEM_DEVICE, MD_MEM_HOST_ASYNC = 0, 1, 2
MD_EGO_FIRST4, MD_EGO_RANSAC_HOMOGRAPHY, MD_EGO_RANSAC_AFFINE = 0, 1, 2
MD_FLOW_LK, MD_FLOW_VARFLOW = 0, 1
STATUS_NAMES = {0: "MD_OK", -1: "MD_ERR_INVALID", -2: "MD_ERR_CUDA", -3: "MD_ERR_NOMEM", -4: "MD_ERR_UNSUPPORTED",
                -5: "MD_ERR_STATE"}

# every symbol include/motion_b200.h declares (tests/test_abi.py checks the library exports all of them)
SYMBOLS = [
    "md_config_default", "md_create", "md_destroy", "md_set_stream", "md_sync", "md_last_error", "md_version",
    "md_grid_size", "md_grid_points", "md_pyramid_levels", "md_pyramid_level_size", "md_gray_u8", "md_pyramid_u8",
    "md_pyramid_read", "md_pyramid_read_deriv", "md_lk_flow", "md_fit_egomotion", "md_motion_mask",
    "md_process_batch", "md_process_pair", "md_track_trajectories", "md_fit_subspace", "md_varflow", "md_stats_get",
    "md_stats_reset", "md_profile", "md_profile_read", "md_live_params_default", "md_window_reset", "md_window_push",
    "md_window_detect", "md_cluster_points", "md_find_outliers", "md_cluster_vectors", "md_draw_flow", "md_set_pair_index",
    "md_unpack_mask_host",
]


class MdConfig(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("max_batch", C.c_int32), ("pixel_step", C.c_int32),
        ("min_vector_size", C.c_double),
        ("lk_win", C.c_int32), ("lk_max_level", C.c_int32), ("lk_max_iters", C.c_int32),
        ("lk_eps", C.c_double), ("lk_min_eig", C.c_float),
        ("diff_threshold", C.c_int32), ("morph", C.c_int32), ("ego_mode", C.c_int32), ("ransac_iters", C.c_int32),
        ("ransac_thresh", C.c_double), ("seed", C.c_uint32),
        ("vf_max_level", C.c_int32), ("vf_start_level", C.c_int32), ("vf_n1", C.c_int32), ("vf_n2", C.c_int32),
        ("vf_rho", C.c_float), ("vf_alpha", C.c_float), ("vf_sigma", C.c_float), ("vf_literal", C.c_int32),
        ("flow_engine", C.c_int32), ("vf_grid_barrier", C.c_int32), ("cuda_graphs", C.c_int32), ("mask_packed", C.c_int32), ("reserved", C.c_int32 * 4),
    ]


class MdFrames(C.Structure):
    _fields_ = [("data", C.c_void_p), ("channels", C.c_int32), ("pitch", C.c_int32), ("frame_stride", C.c_int64),
                ("count", C.c_int32), ("chain", C.c_int32)]


class MdOutputs(C.Structure):
    _fields_ = [("next_pts", C.c_void_p), ("status", C.c_void_p), ("keep", C.c_void_p), ("H", C.c_void_p),
                ("num_vectors", C.c_void_p), ("inliers", C.c_void_p), ("mask", C.c_void_p), ("mask_pitch", C.c_int32),
                ("mask_stride", C.c_int64)]


class MdStats(C.Structure):
    _fields_ = [("pairs", C.c_int64), ("mask_pixels", C.c_int64), ("tracked", C.c_int64), ("inliers", C.c_int64),
                ("last_H", C.c_double * 9), ("kernel_launches", C.c_int64), ("device", C.c_int32), ("reserved0", C.c_int32),
                ("lk_iterations", C.c_int64), ("lk_levels", C.c_int64), ("graph_replays", C.c_int64)]


class MdLiveParams(C.Structure):
    _fields_ = [("num_motions", C.c_int32), ("sigma", C.c_double), ("distance_threshold", C.c_double), ("seed", C.c_uint32),
                ("subspace_iters", C.c_int32), ("min_cluster_size", C.c_int32), ("reserved", C.c_int32 * 4)]


class MdLiveResult(C.Structure):
    _fields_ = [("num_trajectories", C.c_int32), ("subspace_inliers", C.c_int32), ("num_outliers", C.c_int32),
                ("num_clusters_all", C.c_int32), ("num_clusters", C.c_int32), ("reserved", C.c_int32 * 3),
                ("traj", C.c_void_p), ("traj_index", C.c_void_p), ("residual", C.c_void_p), ("outlier", C.c_void_p),
                ("best_cols", C.c_void_p), ("outlier_points", C.c_void_p), ("labels", C.c_void_p), ("boxes", C.c_void_p),
                ("cluster_sizes", C.c_void_p), ("cluster_ids", C.c_void_p)]


class MotionB200Error(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__("%s (%d): %s" % (STATUS_NAMES.get(code, "?"), code, msg))


_lib = None


def lib():
    """Load libmotion_b200.so; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MotionB200Error(-2, "CUDA library %s is missing: build it with __graft_entry__.build(); "
                                  "there is no CPU fallback" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.md_last_error.restype = C.c_char_p
        _lib.md_version.restype = C.c_char_p
        for name in SYMBOLS:
            getattr(_lib, name)
    return _lib


def default_config(**kw):
    cfg = MdConfig()
    lib().md_config_default(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise AttributeError("md_config has no field %r" % k)
        setattr(cfg, k, v)
    return cfg


def _ptr(a):
    """numpy array (host), torch tensor (device or host) or raw int -> void*"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(a.data_ptr())     # torch tensor


def unpack_mask(bits, width):
    """Host utility md_unpack_mask_host: [..., rows, pitch] packed mask bits (LSB first) -> [..., rows, width] bytes 0 / 255."""
    bits = np.ascontiguousarray(bits, np.uint8)
    rows = int(np.prod(bits.shape[:-1]))
    out = np.empty(bits.shape[:-1] + (width,), np.uint8)
    rc = lib().md_unpack_mask_host(_ptr(bits), bits.shape[-1], width, rows, _ptr(out), width)
    if rc != MD_OK:
        raise MotionB200Error(rc, "md_unpack_mask_host")
    return out


class Context:
    """One md_ctx = one camera stream on one GPU (not thread-safe; contexts are independent)."""

    def __init__(self, cfg=None, device=0, **kw):
        self.cfg = cfg if cfg is not None else default_config(**kw)
        self._h = C.c_void_p()
        rc = lib().md_create(C.byref(self.cfg), int(device), C.byref(self._h))
        if rc != MD_OK:
            raise MotionB200Error(rc, "md_create failed (needs a CUDA device with compute capability 10.x)")
        self.device = device
        self.w, self.h = self.cfg.width, self.cfg.height
        self.P = lib().md_grid_size(self._h)
        self.levels = lib().md_pyramid_levels(self._h)

    def close(self):
        if self._h:
            lib().md_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != MD_OK:
            raise MotionB200Error(rc, lib().md_last_error(self._h).decode())

    # ---- plumbing -------------------------------------------------------------------------------------------
    def set_stream(self, cuda_stream_handle):
        self._ck(lib().md_set_stream(self._h, C.c_void_p(cuda_stream_handle)))

    def sync(self):
        self._ck(lib().md_sync(self._h))

    def grid_points(self):
        pts = np.empty((self.P, 2), np.float32)
        lib().md_grid_points(self._h, _ptr(pts))
        return pts

    def level_size(self, level):
        w, h = C.c_int32(), C.c_int32()
        self._ck(lib().md_pyramid_level_size(self._h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    # ---- host-memory (numpy) convenience wrappers; device pointers go through the raw_* methods --------------
    def gray(self, rgb):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        h, w, _ = rgb.shape
        out = np.empty((h, w), np.uint8)
        self._ck(lib().md_gray_u8(self._h, _ptr(rgb), w * 3, w, h, _ptr(out), w, MD_MEM_HOST))
        return out

    def pyramid(self, gray, slot):
        gray = np.ascontiguousarray(gray, np.uint8)
        assert gray.shape == (self.h, self.w)
        self._ck(lib().md_pyramid_u8(self._h, _ptr(gray), self.w, slot, MD_MEM_HOST))

    def pyramid_read(self, slot, level):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        self._ck(lib().md_pyramid_read(self._h, slot, level, _ptr(out), w, MD_MEM_HOST))
        return out

    def pyramid_read_deriv(self, slot, level):
        w, h = self.level_size(level)
        out = np.empty((h, w, 2), np.int16)
        self._ck(lib().md_pyramid_read_deriv(self._h, slot, level, _ptr(out), MD_MEM_HOST))
        return out

    def lk_flow(self, slot_prev, slot_next, pts=None):
        if pts is None:
            n = self.P
            pin = None
        else:
            pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
            n = len(pts)
            pin = _ptr(pts)
        out = np.zeros((n, 2), np.float32)
        st = np.zeros(n, np.uint8)
        self._ck(lib().md_lk_flow(self._h, slot_prev, slot_next, pin, n, _ptr(out), _ptr(st), MD_MEM_HOST))
        return out, st

    def fit_egomotion(self, src, dst, status=None, keep=None, mode=MD_EGO_RANSAC_HOMOGRAPHY, seed=1):
        src = np.ascontiguousarray(src, np.float32).reshape(-1, 2)
        dst = np.ascontiguousarray(dst, np.float32).reshape(-1, 2)
        n = len(src)
        status = None if status is None else np.ascontiguousarray(status, np.uint8)
        keep = None if keep is None else np.ascontiguousarray(keep, np.uint8)
        H = np.zeros(9, np.float64)
        nv, ninl = C.c_int32(), C.c_int32()
        inl = np.zeros(n, np.uint8)
        self._ck(lib().md_fit_egomotion(self._h, _ptr(src), _ptr(dst), _ptr(status), _ptr(keep), n, mode, C.c_uint32(seed),
                                        _ptr(H), C.byref(nv), C.byref(ninl), _ptr(inl), MD_MEM_HOST))
        return dict(H=H.reshape(3, 3), num_vectors=nv.value, inliers=ninl.value, inlier_mask=inl)

    def motion_mask(self, prev, cur, H, thresh=None, morph=None):
        prev = np.ascontiguousarray(prev, np.uint8)
        cur = np.ascontiguousarray(cur, np.uint8)
        H = np.ascontiguousarray(H, np.float64)
        out = np.empty((self.h, self.w), np.uint8)
        thresh = self.cfg.diff_threshold if thresh is None else thresh
        morph = self.cfg.morph if morph is None else int(morph)
        self._ck(lib().md_motion_mask(self._h, _ptr(prev), _ptr(cur), self.w, _ptr(H), thresh, morph, _ptr(out), self.w,
                                      MD_MEM_HOST))
        return out

    def process_batch(self, frames, chain=False, want_mask=True):
        """frames: numpy [n][h][w] (gray) or [n][h][w][3]; returns dict of per-pair numpy outputs."""
        frames = np.ascontiguousarray(frames, np.uint8)
        ch = 3 if frames.ndim == 4 else 1
        n = frames.shape[0]
        pairs = n if chain else n - 1
        fr = MdFrames(_ptr(frames), ch, self.w * ch, self.w * self.h * ch, n, 1 if chain else 0)
        res = dict(next=np.zeros((pairs, self.P, 2), np.float32), status=np.zeros((pairs, self.P), np.uint8),
                   keep=np.zeros((pairs, self.P), np.uint8), H=np.zeros((pairs, 3, 3), np.float64),
                   num_vectors=np.zeros(pairs, np.int32), inliers=np.zeros(pairs, np.int32),
                   mask=np.zeros((pairs, self.h, self.w), np.uint8) if want_mask else None)
        mpitch, mstride = self.w, self.w * self.h
        if want_mask and self.cfg.mask_packed:
            # 1 bit per pixel on the wire; handed back both packed ("mask_bits") and expanded to 0 / 255
            mpitch = (self.w + 7) // 8
            mstride = mpitch * self.h
            res["mask_bits"] = np.zeros((pairs, self.h, mpitch), np.uint8)
        out = MdOutputs(_ptr(res["next"]), _ptr(res["status"]), _ptr(res["keep"]), _ptr(res["H"]), _ptr(res["num_vectors"]),
                        _ptr(res["inliers"]), _ptr(res["mask_bits"] if "mask_bits" in res else res["mask"]), mpitch, mstride)
        self._ck(lib().md_process_batch(self._h, C.byref(fr), C.byref(out), MD_MEM_HOST))
        if "mask_bits" in res:
            res["mask"] = np.unpackbits(res["mask_bits"], axis=2, bitorder="little")[:, :, :self.w] * np.uint8(255)
        return res

    def raw_process_batch(self, frames_ptr, channels, pitch, frame_stride, count, chain, outputs, mem):
        """Thin call for device (or pinned host) pointers. `outputs` is an MdOutputs."""
        fr = MdFrames(C.c_void_p(frames_ptr), channels, pitch, frame_stride, count, 1 if chain else 0)
        self._ck(lib().md_process_batch(self._h, C.byref(fr), C.byref(outputs), mem))

    def set_pair_index(self, index):
        """Number of the next pair (numbers the RANSAC seeds): lets several contexts share one sequence (streams.BatchPipeline)."""
        self._ck(lib().md_set_pair_index(self._h, C.c_uint64(index)))

    def track_trajectories(self, frames):
        frames = np.ascontiguousarray(frames, np.uint8)
        ch = 3 if frames.ndim == 4 else 1
        F = frames.shape[0]
        fr = MdFrames(_ptr(frames), ch, self.w * ch, self.w * self.h * ch, F, 0)
        traj = np.zeros((self.P, F, 2), np.float32)
        ln = np.zeros(self.P, np.int32)
        lp = np.zeros((self.P, 2), np.float32)
        lnx = np.zeros((self.P, 2), np.float32)
        ls = np.zeros(self.P, np.uint8)
        self._ck(lib().md_track_trajectories(self._h, C.byref(fr), _ptr(traj), _ptr(ln), _ptr(lp), _ptr(lnx), _ptr(ls),
                                             MD_MEM_HOST))
        return dict(traj=traj, len=ln, last_prev=lp, last_next=lnx, last_status=ls)

    def fit_subspace(self, traj, num_motions=2, sigma=0.5, seed=1, forced_cols=None, iters=50):
        traj = np.ascontiguousarray(traj, np.float32)
        T, F, _ = traj.shape
        d = 4 * num_motions
        res = np.zeros(T, np.float32)
        cols = np.zeros(d, np.int32)
        outl = np.zeros(T, np.uint8)
        ninl = C.c_int32()
        fc = None
        if forced_cols is not None:
            fc_arr = np.ascontiguousarray(forced_cols, np.int32).reshape(-1, d)
            iters = fc_arr.shape[0]
            fc = _ptr(fc_arr)
        self._ck(lib().md_fit_subspace(self._h, _ptr(traj), T, F, num_motions, C.c_double(sigma), C.c_uint32(seed), fc, iters,
                                       _ptr(res), _ptr(cols), _ptr(outl), C.byref(ninl), MD_MEM_HOST))
        return dict(inliers=ninl.value, residual=res, best_cols=cols, outlier=outl)

    # ---- the node's live path (imageCallback): ring of the last F frames ----------------------------------------
    def window_reset(self):
        self._ck(lib().md_window_reset(self._h))

    def window_push(self, frame):
        frame = np.ascontiguousarray(frame, np.uint8)
        ch = 3 if frame.ndim == 3 else 1
        fill = C.c_int32()
        self._ck(lib().md_window_push(self._h, _ptr(frame), ch, self.w * ch, C.byref(fill), MD_MEM_HOST))
        return fill.value

    def raw_window_push(self, frame_ptr, channels, pitch, mem):
        fill = C.c_int32()
        self._ck(lib().md_window_push(self._h, C.c_void_p(frame_ptr), channels, pitch, C.byref(fill), mem))
        return fill.value

    def raw_window_detect(self, params, result, mem):
        """params: MdLiveParams, result: MdLiveResult with device (or pinned host) pointers / NULLs; counts land in `result`."""
        self._ck(lib().md_window_detect(self._h, C.byref(params), C.byref(result), mem))

    def window_detect(self, num_motions=2, sigma=0.5, distance_threshold=50.0, seed=1, iters=50, min_cluster_size=5):
        lp = MdLiveParams()
        lib().md_live_params_default(C.byref(lp))
        lp.num_motions, lp.sigma, lp.distance_threshold, lp.seed = num_motions, sigma, distance_threshold, seed
        lp.subspace_iters, lp.min_cluster_size = iters, min_cluster_size
        F, P = 2 * num_motions + 1, self.P
        a = dict(traj=np.zeros((P, F, 2), np.float32), traj_index=np.zeros(P, np.int32), residual=np.zeros(P, np.float32),
                 outlier=np.zeros(P, np.uint8), best_cols=np.zeros(4 * num_motions, np.int32),
                 outlier_points=np.zeros((P, 2), np.float32), labels=np.zeros(P, np.int32), boxes=np.zeros((P, 4), np.int32),
                 cluster_sizes=np.zeros(P, np.int32), cluster_ids=np.zeros(P, np.int32))
        r = MdLiveResult()
        for k, v in a.items():
            setattr(r, k, v.ctypes.data)
        self._ck(lib().md_window_detect(self._h, C.byref(lp), C.byref(r), MD_MEM_HOST))
        T, no, K = r.num_trajectories, r.num_outliers, r.num_clusters
        out = dict(num_trajectories=T, subspace_inliers=r.subspace_inliers, num_outliers=no, num_clusters_all=r.num_clusters_all,
                   num_clusters=K)
        out.update(traj=a["traj"][:T], traj_index=a["traj_index"][:T], residual=a["residual"][:T], outlier=a["outlier"][:T],
                   best_cols=a["best_cols"], outlier_points=a["outlier_points"][:no], labels=a["labels"][:no],
                   boxes=a["boxes"][:K], cluster_sizes=a["cluster_sizes"][:K], cluster_ids=a["cluster_ids"][:K])
        return out

    def cluster_points(self, pts, distance_threshold=50.0, min_cluster_size=5):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        n = len(pts)
        labels = np.zeros(max(n, 1), np.int32)
        boxes = np.zeros((max(n, 1), 4), np.int32)
        sizes = np.zeros(max(n, 1), np.int32)
        ids = np.zeros(max(n, 1), np.int32)
        nall, k = C.c_int32(), C.c_int32()
        self._ck(lib().md_cluster_points(self._h, _ptr(pts), n, C.c_double(distance_threshold), min_cluster_size, _ptr(labels),
                                         C.byref(nall), C.byref(k), _ptr(boxes), _ptr(sizes), _ptr(ids), MD_MEM_HOST))
        return labels[:n], nall.value, boxes[:k.value], sizes[:k.value], ids[:k.value]

    def cluster_vectors(self, vec4, distance_threshold=50.0, angular_threshold=0.15):
        """FlowClusterer::getClusters on the participating vectors [n][4] f64: (labels [n], clusters founded)."""
        vec4 = np.ascontiguousarray(vec4, np.float64).reshape(-1, 4)
        n = len(vec4)
        labels = np.zeros(max(n, 1), np.int32)
        nall = C.c_int32()
        self._ck(lib().md_cluster_vectors(self._h, _ptr(vec4), n, C.c_double(distance_threshold), C.c_double(angular_threshold),
                                          _ptr(labels), C.byref(nall), MD_MEM_HOST))
        return labels[:n], nall.value

    def draw_flow(self, image, vec4=None, next_pts=None, status=None, keep=None, colour=(255, 0, 0)):
        """showOpticalFlowVectors: image [h][w] or [h][w][3]; either vec4 [n][4] (row-major field elements) or one pair's
        next_pts / status / keep.  Returns (image with arrows, arrows drawn)."""
        image = np.ascontiguousarray(image, np.uint8)
        ch = 3 if image.ndim == 3 else 1
        out = np.empty_like(image)
        col = np.ascontiguousarray(np.array(colour, np.uint8).reshape(-1)[:ch])
        drawn = C.c_int32()
        if vec4 is not None:
            vec4 = np.ascontiguousarray(vec4, np.float64).reshape(-1, 4)
            vp, n = _ptr(vec4) if len(vec4) else C.c_void_p(8), len(vec4)
            a = b = c = None
        else:
            vp, n = None, 0
            a = np.ascontiguousarray(next_pts, np.float32); b = np.ascontiguousarray(status, np.uint8); c = np.ascontiguousarray(keep, np.uint8)
        self._ck(lib().md_draw_flow(self._h, _ptr(image), ch, self.w * ch, vp, n, _ptr(a), _ptr(b), _ptr(c), _ptr(col), _ptr(out),
                                    self.w * ch, C.byref(drawn), MD_MEM_HOST))
        return out, drawn.value

    def find_outliers(self, dxdy, include_zeros=False):
        dxdy = np.ascontiguousarray(dxdy, np.float64).reshape(-1, 2)
        n = len(dxdy)
        out = np.zeros(n, np.uint8)
        stats = np.zeros(4, np.float64)
        self._ck(lib().md_find_outliers(self._h, _ptr(dxdy), n, 1 if include_zeros else 0, _ptr(out), _ptr(stats), MD_MEM_HOST))
        return out, stats

    def varflow(self, A, B):
        A = np.ascontiguousarray(A, np.uint8)
        B = np.ascontiguousarray(B, np.uint8)
        U = np.zeros((self.h, self.w), np.float32)
        V = np.zeros((self.h, self.w), np.float32)
        self._ck(lib().md_varflow(self._h, _ptr(A), _ptr(B), self.w, _ptr(U), _ptr(V), MD_MEM_HOST))
        return U, V

    def stats(self):
        st = MdStats()
        self._ck(lib().md_stats_get(self._h, C.byref(st)))
        return dict(pairs=st.pairs, mask_pixels=st.mask_pixels, tracked=st.tracked, inliers=st.inliers,
                    last_H=np.array(st.last_H[:]).reshape(3, 3), device=st.device, kernel_launches=st.kernel_launches,
                    lk_iterations=st.lk_iterations, lk_levels=st.lk_levels, graph_replays=st.graph_replays)

    def profile(self, enable=True):
        self._ck(lib().md_profile(self._h, 1 if enable else 0))

    def profile_read(self):
        """ms of the four stages (pyramid, lk, egomotion, mask) of the last profiled md_process_batch."""
        ms = (C.c_float * 4)()
        self._ck(lib().md_profile_read(self._h, ms))
        return [float(v) for v in ms]

    def stats_reset(self):
        self._ck(lib().md_stats_reset(self._h))
