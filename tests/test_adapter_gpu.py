"""GPU: the reference-named C++ classes of adapter/ (OpticalFlowCalculator, OutlierDetector, VarFlow) driven end to end
by adapter/adapter_test, compared with the oracle -- "the node keeps calling the same methods"."""
import os
import subprocess

import numpy as np
import pytest

from motion_detection_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_adapter_classes_end_to_end(oracle, tmp_path):
    exe = os.path.join(ROOT, "adapter", "adapter_test")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "adapter"), "-s"])
    w, h, F = 320, 240, 5
    frames, _ = synth.sequence(w, h, F, seed=1234, blobs=2)
    fin, fout = tmp_path / "frames.bin", tmp_path / "out.bin"
    frames.tofile(fin)
    r = subprocess.run([exe, str(fin), str(w), str(h), str(F), str(fout)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    buf = open(fout, "rb").read()
    off = 0

    def take(dtype, n):
        nonlocal off
        a = np.frombuffer(buf, dtype=dtype, count=n, offset=off)
        off += a.nbytes
        return a

    nv, crows, ccols, ftype = take(np.int32, 4)
    H = take(np.float64, 9).reshape(3, 3)
    # `comp` is the thresholded difference as at cpp:124-127: no erode / dilate unless setMorphology(true)
    ref = oracle.process_pair(frames[0], frames[1], pixel_step=10, min_vector_size=0.2, seed=11, morph=False)
    assert abs(int(nv) - ref["num_vectors"]) <= 2 and nv > 500
    assert (crows, ccols) == (h, w)
    assert ftype == 6 + (3 << 3)                       # CV_64FC4: the adapter repairs the node's CV_32FC4 allocation
    assert np.linalg.norm(H - ref["H"]) / np.linalg.norm(ref["H"]) < 1e-4
    mask = take(np.uint8, w * h).reshape(h, w)
    assert (mask == ref["mask"]).mean() >= 0.999
    P = len(ref["pts"])
    flow = take(np.float64, 4 * P).reshape(P, 4)
    ok = (ref["keep"] == 1) & (flow[:, 0] >= 0)
    assert np.array_equal(flow[ok, :2], ref["pts"][ok].astype(np.float64))          # grid order x outer / y inner (cpp:56-64)
    assert np.abs(flow[ok, 2:] - (ref["next"][ok] - ref["pts"][ok])).mean() < 0.01
    failed = ref["status"] == 0
    if failed.any():
        assert (flow[failed, 0] == -1).mean() > 0.9                                   # failed points: (-1,-1,0,0), cpp:112-115

    # calculateCompensatedFlow: LK with maxLevel 2, vectors over 1 px
    nvc = int(take(np.int32, 1)[0])
    flowc = take(np.float64, 4 * P).reshape(P, 4)
    refc, stc = oracle.lk(frames[0], frames[1], ref["pts"], max_level=2)
    dc = refc - ref["pts"]
    keepc = (stc == 1) & ((np.abs(dc[:, 0]) > 1.0) | (np.abs(dc[:, 1]) > 1.0))
    assert abs(nvc - int(keepc.sum())) <= 2
    okc = keepc & (flowc[:, 0] >= 0) & ((flowc[:, 2] != 0) | (flowc[:, 3] != 0))
    assert okc.sum() >= keepc.sum() - 2 and np.abs(flowc[okc, 2:] - dc[okc]).mean() < 0.01

    # findOutliers over the adapter == the oracle's createMask on the same (dx, dy) field (row-major grid traversal)
    nflag, nvec = take(np.int32, 2)
    g = flow.reshape(w // 10, h // 10, 4).transpose(1, 0, 2).reshape(-1, 4)
    mad_ref, _ = oracle.find_outliers(g[:, 2:4], False)
    assert abs(int(nflag) - int(mad_ref.sum())) <= 1 and nvec <= nflag

    nv2, ntraj, nout, nbasis, ninl = take(np.int32, 5)
    traj = take(np.float32, int(ntraj) * F * 2).reshape(int(ntraj), F, 2)
    assert ntraj > 0.7 * P and nbasis == 8 and 0 < ninl <= ntraj
    n, res, cols, outl, thr = oracle.fit_subspace(traj, num_motions=2, sigma=0.5, seed=3)
    assert n == ninl and int(outl.sum()) == nout

    # FlowClusterer::clusterEuclidean over the adapter == the oracle's literal loop on the same outlier points
    nc = int(take(np.int32, 1)[0])
    got_clusters = []
    for _ in range(nc):
        sz = int(take(np.int32, 1)[0])
        box = take(np.int32, 4).copy()
        got_clusters.append((sz, box, take(np.float32, 2 * sz).reshape(sz, 2).copy()))
    opts = take(np.float32, 2 * int(nout)).reshape(int(nout), 2)
    lab, nall, boxes, sizes, ids = oracle.cluster_euclidean(opts, 50.0, 5)
    assert nc == len(boxes)
    for c, (sz, box, members) in enumerate(got_clusters):
        assert sz == sizes[c] and np.array_equal(box, boxes[c])
        assert np.array_equal(members, opts[lab == ids[c]])

    vf_ok = take(np.int32, 1)[0]
    U = take(np.float32, w * h).reshape(h, w)
    V = take(np.float32, w * h).reshape(h, w)
    assert vf_ok == 1 and off == len(buf)
    Uo, Vo = oracle.varflow(frames[0], frames[1])
    assert np.sqrt((U - Uo) ** 2 + (V - Vo) ** 2).mean() < 0.01


def test_node_call_sites_run_on_the_device(tmp_path):
    """adapter/test/node_callsites.cpp with "gpu": every ofc_. / od_. / fc_. call expression of the node executed once."""
    exe = os.path.join(ROOT, "adapter", "node_callsites")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "adapter"), "-s"])
    r = subprocess.run([exe, "gpu", str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    last = [l for l in r.stdout.splitlines() if l.startswith("== gpu")]
    assert last, r.stdout
    ntraj, ftype = int(last[0].split()[3]), int(last[0].split()[-1])
    assert ntraj > 300 and ftype == 6 + (3 << 3)
