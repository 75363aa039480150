"""GPU: the chain with the dense variational engine as the flow source (BASELINE configs[2]: "dense flow + homography
egomotion"): VarFlow::CalcFlow -> grid sampling -> vector filter -> RANSAC homography -> fused mask, against the oracle."""
import numpy as np
import pytest

from motion_detection_b200 import synth

pytestmark = pytest.mark.gpu


def _slow_camera_sequence(w, h, n, seed):
    # VarFlow has no warping step: it is accurate for sub-pixel motion only, so the camera moves slowly
    canvas = synth.texture(w, h, seed, margin=64, lo=0.0, hi=200.0)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    frames = np.empty((n, h, w), np.uint8)
    for k in range(n):
        frames[k] = np.clip(np.rint(synth._sample_bilinear(canvas, xx + 64 + 0.45 * k, yy + 64 - 0.3 * k)), 0, 255)
    return frames


@pytest.mark.parametrize("size", [(320, 240), (640, 480)])
def test_varflow_engine_chain_matches_oracle(capi, oracle, size):
    w, h = size
    frames = _slow_camera_sequence(w, h, 3, seed=21)
    ctx = capi.Context(width=w, height=h, max_batch=2, pixel_step=10, min_vector_size=0.1, seed=9,
                       flow_engine=capi.MD_FLOW_VARFLOW)
    res = ctx.process_batch(frames)
    for p in range(2):
        ref = oracle.process_pair_varflow(frames[p], frames[p + 1], pixel_step=10, min_vector_size=0.1, seed=9 + p)
        assert (res["status"][p] == 1).all()
        assert np.linalg.norm(res["next"][p] - ref["next"], axis=1).mean() < 0.01
        assert abs(int(res["num_vectors"][p]) - ref["num_vectors"]) <= 0.002 * len(ref["pts"]) + 1
        assert np.linalg.norm(res["H"][p] - ref["H"]) / np.linalg.norm(ref["H"]) < 1e-4
        assert (res["mask"][p] == ref["mask"]).mean() >= 0.999
        # the estimated egomotion is the (sub-pixel) camera translation, underestimated like the reference's VarFlow does
        assert 0.1 < -res["H"][p][0, 2] < 0.6 and 0.05 < res["H"][p][1, 2] < 0.45
