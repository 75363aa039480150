"""The reference's chain restated with the SAME native OpenCV routines the reference calls, via cv2 (4.13.0 in
this image).  This is the PIN for oracle/: tests compare the plain-C oracle with these functions live, and
tests/golden/make_golden.py freezes their outputs into fixtures.  Test infrastructure only.

  calculate_optical_flow   OpticalFlowCalculator::calculateOpticalFlow  common/src/optical_flow_calculator.cpp:30-130
  morph_open               BackgroundSubtractor erode/dilate            common/src/background_subtractor.cpp:31-32
  varflow                  VarFlow::CalcFlow                             common/src/VarFlow.cpp:600-697
"""
import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

LK_KW = dict(winSize=(40, 40), maxLevel=5, flags=0, minEigThreshold=0.001)


def have_cv2():
    return cv2 is not None


def lk(prev, cur, pts):
    crit = (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 10, 0.03)
    p2, st, _ = cv2.calcOpticalFlowPyrLK(prev, cur, np.ascontiguousarray(pts, np.float32).reshape(-1, 1, 2), None,
                                         criteria=crit, **LK_KW)
    return p2.reshape(-1, 2), st.ravel()


def grid(w, h, ps):
    return np.array([(i, j) for i in range(0, w, ps) for j in range(0, h, ps)], np.float32)


def mask_chain(prev, cur, H, thresh=190, morph=True):
    comp = cv2.warpPerspective(prev, H, (prev.shape[1], prev.shape[0]))
    d = cv2.absdiff(comp, cur)
    _, m = cv2.threshold(d, thresh, 255, cv2.THRESH_BINARY)
    if morph:
        m = cv2.erode(m, None)
        m = cv2.dilate(m, None)
    return m


def calculate_optical_flow_strict(prev, cur, pixel_step, min_vector_size):
    """Literal calculateOpticalFlow: H from the first four surviving vectors (degenerate for column x=0)."""
    h, w = prev.shape
    pts = grid(w, h, pixel_step)
    p2, st = lk(prev, cur, pts)
    d = p2 - pts
    keep = (st == 1) & ((np.abs(d[:, 0]) > min_vector_size) | (np.abs(d[:, 1]) > min_vector_size))
    src, dst = pts[keep], p2[keep]
    H = None
    if keep.sum() >= 4:
        H = cv2.getPerspectiveTransform(src[:4], dst[:4])
    return pts, p2, st, keep, H


# ------------------------------------------------------------------------------------------------------
# VarFlow with cv2 primitives (legacy C API semantics: BORDER_REPLICATE, correlation) + plain python GS
# ------------------------------------------------------------------------------------------------------

def _gs_iteration(U, V, J11, J12, J13, J22, J23, h, alpha, iters):
    hh, ww = U.shape
    f = np.float32
    c = f(f(h) * f(h) / f(alpha))
    for _ in range(iters):
        for y in range(hh):
            for x in range(ww):
                n = 0
                t = f(0)
                if y - 1 > -1:
                    t = f(t + U[y - 1, x]); n += 1
                if y + 1 < hh:
                    t = f(t + U[y + 1, x]); n += 1
                if x - 1 > -1:
                    t = f(t + U[y, x - 1]); n += 1
                if x + 1 < ww:
                    t = f(t + U[y, x + 1]); n += 1
                t = f(t - f(c * f(f(J12[y, x] * V[y, x]) + J13[y, x])))
                U[y, x] = f(t / f(f(n) + f(c * J11[y, x])))
                n = 0
                t = f(0)
                if y - 1 > -1:
                    t = f(t + V[y - 1, x]); n += 1
                if y + 1 < hh:
                    t = f(t + V[y + 1, x]); n += 1
                if x - 1 > -1:
                    t = f(t + V[y, x - 1]); n += 1
                if x + 1 < ww:
                    t = f(t + V[y, x + 1]); n += 1
                t = f(t - f(c * f(f(J12[y, x] * U[y, x]) + J23[y, x])))
                V[y, x] = f(t / f(f(n) + f(c * J22[y, x])))


def _residual(U, V, J11, J12, J22, J13, J23, Ur, Vr, h, alpha):
    hh, ww = U.shape
    f = np.float32
    ih2 = f(f(1) / f(f(h) * f(h)))
    ia = f(f(1) / f(alpha))
    ur = np.empty_like(U)
    vr = np.empty_like(V)
    for (F_, G_, Jd, out) in ((U, V, J11, ur), (V, U, J22, vr)):
        Pd = np.pad(F_, 1)
        nb = Pd[:-2, 1:-1]
        # neighbour sum in the reference's order: top, bottom, left, right, skipping missing ones
        t = np.zeros_like(F_)
        n = np.zeros(F_.shape, np.float32)
        ys, xs = np.mgrid[0:hh, 0:ww]
        m = ys > 0
        t[m] = (t + Pd[:-2, 1:-1])[m]; n[m] += 1
        m = ys < hh - 1
        t[m] = (t + Pd[2:, 1:-1])[m]; n[m] += 1
        m = xs > 0
        t[m] = (t + Pd[1:-1, :-2])[m]; n[m] += 1
        m = xs < ww - 1
        t[m] = (t + Pd[1:-1, 2:])[m]; n[m] += 1
        r = (n * F_ - t).astype(f)
        r = (r * ih2).astype(f)
        r = (r - (ia * (Jd * F_ + J12 * G_).astype(f)).astype(f)).astype(f)
        out[...] = r
    # literal order (VarFlow.cpp:476-489): the part-residual is stored first, THEN combined with J13 -- which at
    # recursion depth >= 1 is the very same buffer (VarFlow.cpp:537), i.e. already overwritten.
    Ur[...] = ur
    Vr[...] = vr
    Ur[...] = (J13 * ia + Ur * f(-1)).astype(f)
    Vr[...] = (J23 * ia + Vr * f(-1)).astype(f)


def varflow(A, B, max_level=4, n1=2, n2=2, rho=2.8, alpha=1400.0, sigma=1.5, literal=True, gs=None):
    """gs: optional replacement for the python GS sweep (e.g. a numba-jitted copy) with the same signature."""
    gs = gs or _gs_iteration
    h, w = A.shape
    f = np.float32

    def blur(x, s):
        return cv2.GaussianBlur(x, (0, 0), s, borderType=cv2.BORDER_REPLICATE)

    def resize(x, size):
        return cv2.resize(x, size, interpolation=cv2.INTER_LINEAR)

    Af = blur(A.astype(f), sigma)
    Bf = blur(B.astype(f), sigma)
    mx = np.array([[0.08333, -0.66666, 0, 0.66666, -0.08333]], f)
    my = np.array([[-0.08333], [0.66666], [0], [-0.66666], [0.08333]], f)
    fx = cv2.filter2D(Af, -1, mx, borderType=cv2.BORDER_REPLICATE)
    fy = cv2.filter2D(Af, -1, my, borderType=cv2.BORDER_REPLICATE)
    ft = Bf - Af
    nl = max_level + 1
    sizes = [(int(np.floor(w / 2.0 ** i)), int(np.floor(h / 2.0 ** i))) for i in range(nl)]
    J = {k: [blur(v, rho)] for k, v in dict(J11=fx * fx, J12=fx * fy, J13=fx * ft, J22=fy * fy, J23=fy * ft).items()}
    for i in range(1, nl):
        for k in J:
            J[k].append(resize(J[k][i - 1], sizes[i]))
    U = [np.zeros((s[1], s[0]), f) for s in sizes]
    V = [np.zeros((s[1], s[0]), f) for s in sizes]
    Ur = [np.zeros((s[1], s[0]), f) for s in sizes]
    Vr = [np.zeros((s[1], s[0]), f) for s in sizes]

    def rec(lvl, hgrid, J13a, J23a):
        def it(n):
            gs(U[lvl], V[lvl], J["J11"][lvl], J["J12"][lvl], J13a[lvl], J["J22"][lvl], J23a[lvl], hgrid, alpha, n)
        if lvl == max_level:
            it(n1)
            return
        it(n1)
        for cyc in range(2):
            if literal:
                _residual(U[lvl], V[lvl], J["J11"][lvl], J["J12"][lvl], J["J22"][lvl], J13a[lvl], J23a[lvl], Ur[lvl], Vr[lvl],
                          hgrid, alpha)
                Ur[lvl + 1][...] = resize(Ur[lvl], sizes[lvl + 1])
                Vr[lvl + 1][...] = resize(Vr[lvl], sizes[lvl + 1])
                U[lvl + 1][...] = 0
                V[lvl + 1][...] = 0
                rec(lvl + 1, 2 * hgrid, Ur, Vr)
                Ur[lvl][...] = resize(U[lvl + 1], sizes[lvl])
                Vr[lvl][...] = resize(V[lvl + 1], sizes[lvl])
                U[lvl] += Ur[lvl]
                V[lvl] += Vr[lvl]
            it(n1 + n2 if cyc == 0 else n2)

    k = max_level
    while True:
        rec(k, float(2 ** k), J["J13"], J["J23"])
        if k > 0:
            U[k - 1][...] = resize(U[k], sizes[k - 1])
            V[k - 1][...] = resize(V[k], sizes[k - 1])
            k -= 1
        else:
            break
    return U[0], V[0]


def show_optical_flow_vectors(image, vec4, pixel_step, min_vector_size, colour):
    """OpticalFlowVisualizer::showOpticalFlowVectors (common/src/optical_flow_visualizer.cpp:23-71) with cv2.line itself; vec4 =
    the field's elements in the order the reference's double loop meets them."""
    out = image.copy()
    f32 = np.float32
    n = 0
    for e in np.asarray(vec4, np.float64).reshape(-1, 4):
        if not ((abs(e[2]) > min_vector_size or abs(e[3]) > min_vector_size) and abs(e[2]) < pixel_step * 5 and abs(e[3]) < pixel_step * 5):
            continue
        sx, sy = f32(e[0]), f32(e[1])
        ex, ey = f32(np.float64(sx) + e[2]), f32(np.float64(sy) + e[3])
        back = np.float64(np.arctan2(f32(sy - ey), f32(sx - ex), dtype=f32))          # atan2 on float arguments
        pts = [(sx, sy), (ex, ey)]
        for ang in (back + np.pi / 4.0, back - np.pi / 4.0):
            pts.append((f32(np.float64(ex) + 3.0 * np.cos(ang)), f32(np.float64(ey) + 3.0 * np.sin(ang))))
        ip = [(int(np.rint(p[0])), int(np.rint(p[1]))) for p in pts]                   # Point2f -> Point: cvRound
        cv2.line(out, ip[0], ip[1], colour, 1, cv2.LINE_AA, 0)
        cv2.line(out, ip[1], ip[2], colour, 1, cv2.LINE_AA, 0)
        cv2.line(out, ip[1], ip[3], colour, 1, cv2.LINE_AA, 0)
        n += 1
    return out, n
