"""The reference's CPU path for the motion-mask chain, timed by bench.py (cpu_baseline / --impl reference).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference's C++ cannot be compiled in this image (no OpenCV / Eigen /
ROS headers), so the chain of OpenCV calls that OpticalFlowCalculator::calculateOpticalFlow makes
(common/src/optical_flow_calculator.cpp:30-130) plus the erode/dilate of BackgroundSubtractor
(common/src/background_subtractor.cpp:31-32) is issued through cv2 -- the same native OpenCV routines, all host
threads -- when cv2 is importable, and through the plain-C oracle (one thread per point chunk) otherwise.
The egomotion fit between LK and the warp is the oracle's composition; it is < 1 % of the time.
"""
import numpy as np

from . import oracle as O

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def have_cv2():
    return cv2 is not None


def threads():
    import os
    return cv2.getNumThreads() if cv2 is not None else min(16, os.cpu_count() or 1)


def process_pair(prev, cur, pixel_step=10, min_vector_size=0.2, seed=1, mode=O.MODE_RANSAC_HOMOGRAPHY):
    h, w = prev.shape
    pts = O.grid_points(w, h, pixel_step)
    if cv2 is not None:
        crit = (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 10, 0.03)
        p2, st, _ = cv2.calcOpticalFlowPyrLK(prev, cur, pts.reshape(-1, 1, 2), None, winSize=(40, 40), maxLevel=5,
                                             criteria=crit, flags=0, minEigThreshold=0.001)
        p2 = p2.reshape(-1, 2)
        st = st.ravel()
    else:
        p2, st = O.lk(prev, cur, pts)
    nv, keep, _ = O.flow_filter(pts, p2, st, min_vector_size)
    ninl, H, _ = O.fit_egomotion(pts, p2, keep, w, h, mode, 50, 0.5, seed)
    if ninl == 0:
        return np.zeros_like(prev), H, nv
    if cv2 is not None:
        comp = cv2.warpPerspective(prev, H, (w, h))
        d = cv2.absdiff(comp, cur)
        _, m = cv2.threshold(d, 190, 255, cv2.THRESH_BINARY)
        m = cv2.erode(m, None)
        m = cv2.dilate(m, None)
    else:
        m = O.motion_mask(prev, cur, H)
    return m, H, nv


def live_callback(window, pixel_step=10, num_motions=2, sigma=0.5, distance_threshold=50.0, seed=1):
    """One imageCallback body of the reference (node.cpp:262-395) over a window of F frames, the way the reference runs
    it: EVERY frame of the window is converted and pyramided again, the grid is re-tracked through all F - 1 pairs
    (calculateOpticalFlowTrajectory, cpp:133-257, OpenCV calls through cv2), then fitSubspace + clusterEuclidean +
    boundingRect from the oracle's restatement."""
    F = len(window)
    gray = []
    for im in window:
        if im.ndim == 3:
            gray.append(cv2.cvtColor(im, cv2.COLOR_BGR2GRAY) if cv2 is not None else O.gray(im))
        else:
            gray.append(im)
    h, w = gray[0].shape
    pts = O.grid_points(w, h, pixel_step)
    P = len(pts)
    cur = pts.copy()
    traj = np.zeros((P, F, 2), np.float32)
    traj[:, 0] = pts
    ln = np.ones(P, np.int32)
    crit = (3, 10, 0.03)
    for j in range(F - 1):
        if cv2 is not None:
            _, pyr = cv2.buildOpticalFlowPyramid(gray[j], (40, 40), 5, None, True)      # cpp:170 (result unused by the binding)
            p2, st, _ = cv2.calcOpticalFlowPyrLK(gray[j], gray[j + 1], cur.reshape(-1, 1, 2), None, winSize=(40, 40), maxLevel=5,
                                                 criteria=crit, flags=0, minEigThreshold=0.001)
            p2 = np.ascontiguousarray(p2.reshape(-1, 2), np.float32)
            st = np.ascontiguousarray(st.ravel(), np.uint8)
        else:
            p2, st = O.lk(gray[j], gray[j + 1], cur)
        O.lib().orc_traj_step(cur.ctypes.data_as(O.f32p), p2.ctypes.data_as(O.f32p), st.ctypes.data_as(O.u8p),
                              traj.ctypes.data_as(O.f32p), ln.ctypes.data_as(O.i32p), P, F, w, h)
    idx = np.nonzero(ln == F)[0]
    tc = np.ascontiguousarray(traj[idx])
    if len(idx) == 0:
        return dict(num_trajectories=0, boxes=np.zeros((0, 4), np.int32))
    n, res, cols, outl, thr = O.fit_subspace(tc, num_motions=num_motions, sigma=sigma, seed=seed)
    opts = np.ascontiguousarray(tc[outl != 0][:, F - 2])
    labels, ncl, boxes, sizes, ids = O.cluster_euclidean(opts, distance_threshold, 5)
    return dict(num_trajectories=len(idx), num_outliers=len(opts), boxes=boxes, cluster_sizes=sizes)
