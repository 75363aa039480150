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
