"""The CUDA chain against outputs of the REFERENCE's own code (tests/golden/golden_ref.npz, written by
tests/golden/make_golden_ref.py from oracle/_ref/libofc_ref.so = the unmodified common/src/optical_flow_calculator.cpp).
No oracle in between: what the device returns in its literal mode (first-four getPerspectiveTransform, no morphology) is
compared with what OpticalFlowCalculator::calculateOpticalFlow / calculateOpticalFlowTrajectory returned for the same frames."""
import os

import numpy as np
import pytest

from motion_detection_b200 import synth
from test_oracle_ref import _mover_pair

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_ref.npz")


def test_device_chain_equals_reference_code_outputs(capi):
    G = np.load(GOLD)
    f0, f1, ps = _mover_pair()
    h, w = f0.shape
    ctx = capi.Context(width=w, height=h, max_batch=1, pixel_step=ps, min_vector_size=1.0, ego_mode=capi.MD_EGO_FIRST4,
                       morph=0, seed=1)
    for frames in (np.stack([f0, f1]), np.repeat(np.stack([f0, f1])[..., None], 3, axis=3)):      # gray and BGR input
        r = ctx.process_batch(frames)
        assert int(r["num_vectors"][0]) == int(G["ofc_nv"][0])
        grid = G["ofc_flow_grid"]                                   # [gy][gx][4] Vec4d elements of the reference's field
        gx, gy = grid.shape[1], grid.shape[0]
        nxt = r["next"][0].reshape(gx, gy, 2)                       # x outer, y inner (cpp:56-64)
        st = r["status"][0].reshape(gx, gy)
        keep = r["keep"][0].reshape(gx, gy)
        for ix in range(gx):
            for iy in range(gy):
                e = grid[iy, ix]
                if e[0] == -1.0:
                    assert st[ix, iy] == 0
                    continue
                assert st[ix, iy] == 1 and e[0] == ix * ps and e[1] == iy * ps
                moved = e[2] != 0.0 or e[3] != 0.0
                assert bool(keep[ix, iy]) == moved
                if moved:
                    assert abs(nxt[ix, iy, 0] - (e[0] + e[2])) < 0.01 and abs(nxt[ix, iy, 1] - (e[1] + e[3])) < 0.01
        assert int(r["inliers"][0]) == 4
        ref_mask = np.unpackbits(G["ofc_comp_bits"])[: h * w].reshape(h, w) * np.uint8(255)
        assert (r["mask"][0] == ref_mask).mean() >= 0.999 and int(ref_mask.sum()) > 0
    ctx.close()


def test_device_trajectories_equal_reference_code_outputs(capi):
    G = np.load(GOLD)
    frames, _ = synth.sequence(320, 240, 5, seed=21)
    ctx = capi.Context(width=320, height=240, max_batch=4, pixel_step=20, min_vector_size=1.0, seed=1)
    r = ctx.track_trajectories(frames)
    traj, ln = r["traj"], r["len"]
    full = traj[ln == 5]
    ref = G["traj_complete"]
    assert full.shape == ref.shape                                   # the same trajectories survive, in the same (grid) order
    assert np.array_equal(full[:, 0], ref[:, 0])
    assert np.abs(full - ref).max() < 0.05 and np.abs(full - ref).mean() < 0.002
    ctx.close()


def test_device_outlier_detector_equals_reference_code_outputs(capi):
    """md_find_outliers and md_fit_subspace against the frozen outputs of the reference's own outlier_detector.cpp"""
    from test_oracle_ref import _mad_field, _two_motion_trajectories
    G = np.load(GOLD)
    ctx = capi.Context(width=320, height=240, max_batch=1, pixel_step=10, seed=1)
    nodes = _mad_field(3, ps=10, zero_frac=0.3)[::10, ::10].reshape(-1, 4)
    for inc, key in ((False, "mad_flags_nozero"), (True, "mad_flags_zero")):
        r = ctx.find_outliers(nodes[:, 2:4], include_zeros=inc)
        flags = r["outlier"] if isinstance(r, dict) else r[0]
        assert np.array_equal(np.asarray(flags).astype(np.uint8), G[key])
    traj, _ = _two_motion_trajectories(7)
    r = ctx.fit_subspace(traj, num_motions=2, sigma=0.5, seed=7)
    cols = r["best_cols"] if isinstance(r, dict) else r[2]
    outl = r["outlier"] if isinstance(r, dict) else r[3]
    assert np.array_equal(np.asarray(cols), G["sub_cols"]) and np.array_equal(np.nonzero(np.asarray(outl))[0], G["sub_outliers"])
    ctx.close()
