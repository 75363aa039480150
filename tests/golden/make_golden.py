"""Generate tests/golden/golden_cv2.npz: outputs of the OpenCV routines the reference calls (cv2 4.13.0, the
library pin of this build -- the reference itself ships no golden vectors) on small seeded
inputs.  The plain-C oracle is checked against these fixtures in tests/test_oracle_golden.py, so the pin also holds
on hosts without cv2.   Run:  python tests/golden/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import cvref  # noqa: E402
from motion_detection_b200 import synth  # noqa: E402


def main():
    out = {}
    rng = np.random.default_rng(20240607)
    # a2: cvtColor(BGR2GRAY)
    rgb = rng.integers(0, 256, (24, 36, 3), dtype=np.uint8)
    out["gray_in"] = rgb
    out["gray_out"] = cv2.cvtColor(rgb, cv2.COLOR_BGR2GRAY)
    # a3: buildOpticalFlowPyramid(winSize 40, maxLevel 5, withDerivatives)
    img = rng.integers(0, 256, (123, 171), dtype=np.uint8)
    lv, pyr = cv2.buildOpticalFlowPyramid(img, (40, 40), 5, None, True)
    out["pyr_in"] = img
    out["pyr_levels"] = np.int32(lv + 1)
    for l in range(lv + 1):
        out["pyr_l%d" % l] = np.ascontiguousarray(pyr[2 * l])
        out["pyr_d%d" % l] = np.ascontiguousarray(pyr[2 * l + 1])
    # a4: calcOpticalFlowPyrLK on a synthetic pair
    frames, Hs = synth.sequence(320, 240, 2, seed=77, blobs=2)
    pts = cvref.grid(320, 240, 16)
    p2, st = cvref.lk(frames[0], frames[1], pts)
    out["lk_f0"], out["lk_f1"], out["lk_pts"], out["lk_next"], out["lk_status"] = frames[0], frames[1], pts, p2, st
    # a6: getPerspectiveTransform on a non-degenerate quadruple
    src = np.array([[10, 20], [300, 15], [290, 220], [25, 230]], np.float32)
    Ht = np.array([[1.02, 0.03, 2.5], [-0.02, 0.98, -1.75], [2e-5, -3e-5, 1.0]])
    q = (Ht @ np.c_[src, np.ones(4)].T).T
    dst = (q[:, :2] / q[:, 2:]).astype(np.float32)
    out["p4_src"], out["p4_dst"], out["p4_H"] = src, dst, cv2.getPerspectiveTransform(src, dst)
    # a7-a9: warpPerspective / absdiff / threshold / erode / dilate
    a = rng.integers(0, 256, (97, 131), dtype=np.uint8)
    b = rng.integers(0, 256, (97, 131), dtype=np.uint8)
    Hl = [np.eye(3), np.array([[1, 0, 1 / 64.0], [0, 1, 0.5], [0, 0, 1.0]]), Ht,
          np.array([[1.1, 0.2, -30.3], [-0.1, 0.9, 20.7], [1e-4, -2e-4, 1.0]])]
    out["mask_a"], out["mask_b"], out["mask_H"] = a, b, np.array(Hl)
    out["mask_warp"] = np.array([cv2.warpPerspective(a, H, (131, 97)) for H in Hl])
    out["mask_t40"] = np.array([cvref.mask_chain(a, b, H, thresh=40) for H in Hl])
    out["mask_t40_nomorph"] = np.array([cvref.mask_chain(a, b, H, thresh=40, morph=False) for H in Hl])
    out["mask_chain"] = np.array([cvref.mask_chain(frames[0], frames[1], H) for H in (np.eye(3), Hs[0])])
    # a11-a14: VarFlow built from cv2 primitives (GaussianBlur / filter2D / resize) + literal Gauss-Seidel
    vf, _ = synth.sequence(64, 48, 2, seed=5, camera=False, blobs=0, whole_field=(0.75, -0.5), margin=16)
    U, V = cvref.varflow(vf[0], vf[1])
    out["vf_a"], out["vf_b"], out["vf_U"], out["vf_V"] = vf[0], vf[1], U, V
    x = (rng.random((33, 47)) * 100).astype(np.float32)
    out["prim_in"] = x
    out["prim_blur15"] = cv2.GaussianBlur(x, (0, 0), 1.5, borderType=cv2.BORDER_REPLICATE)
    out["prim_blur28"] = cv2.GaussianBlur(x, (0, 0), 2.8, borderType=cv2.BORDER_REPLICATE)
    out["prim_resize_23_16"] = cv2.resize(x, (23, 16), interpolation=cv2.INTER_LINEAR)
    out["prim_resize_94_66"] = cv2.resize(x, (94, 66), interpolation=cv2.INTER_LINEAR)
    x2 = (rng.random((32, 48)) * 100).astype(np.float32)
    out["prim_in2"] = x2
    out["prim_resize_half"] = cv2.resize(x2, (24, 16), interpolation=cv2.INTER_LINEAR)
    # cv::line(..., 1, CV_AA): random segments (some clipped by the image), 3 channels and 1 channel, drawn one after another
    rng = np.random.default_rng(77)
    img3 = rng.integers(0, 256, (60, 80, 3), dtype=np.uint8)
    img1 = rng.integers(0, 256, (60, 80), dtype=np.uint8)
    segs = np.stack([rng.integers(-20, 100, 40), rng.integers(-20, 80, 40), rng.integers(-20, 100, 40), rng.integers(-20, 80, 40)], axis=1).astype(np.int32)
    cols = rng.integers(0, 256, (40, 3), dtype=np.uint8)
    out["aa_img3"], out["aa_img1"], out["aa_segs"], out["aa_cols"] = img3.copy(), img1.copy(), segs, cols
    for s_, c_ in zip(segs, cols):
        cv2.line(img3, (int(s_[0]), int(s_[1])), (int(s_[2]), int(s_[3])), tuple(int(v) for v in c_), 1, cv2.LINE_AA, 0)
        cv2.line(img1, (int(s_[0]), int(s_[1])), (int(s_[2]), int(s_[3])), (int(c_[0]),), 1, cv2.LINE_AA, 0)
    out["aa_out3"], out["aa_out1"] = img3, img1
    # showOpticalFlowVectors on a small field (arrows long enough to overlap their neighbours)
    fimg = rng.integers(0, 256, (90, 120, 3), dtype=np.uint8)
    gy, gx = np.mgrid[0:90:10, 0:120:10]
    vec = np.stack([gx.ravel(), gy.ravel(), rng.normal(6, 9, gx.size), rng.normal(-3, 9, gx.size)], axis=1).astype(np.float64)
    vec[::7, 2:] = 0.0
    out["flowdraw_img"], out["flowdraw_vec"] = fimg, vec
    out["flowdraw_out"], nd = cvref.show_optical_flow_vectors(fimg, vec, 10, 0.2, (255, 0, 0))
    out["flowdraw_n"] = np.int32(nd)
    out["cv2_version"] = np.array(cv2.__version__)
    path = os.path.join(HERE, "golden_cv2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
