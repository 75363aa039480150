"""GPU: md_draw_flow (OpticalFlowVisualizer::showOpticalFlowVectors on the device, k_draw.cu) against the oracle's sequential
cv::line loop: bit-identical images -- anti-aliased arrows blended in the reference's drawing order -- from a Vec4d list and
straight from a batch's next_pts / status / keep."""
import numpy as np
import pytest

from motion_detection_b200 import synth

pytestmark = pytest.mark.gpu


def _field(rng, w, h, ps, mean, sd):
    gy, gx = np.mgrid[0:h:ps, 0:w:ps]
    vec = np.stack([gx.ravel(), gy.ravel(), rng.normal(mean[0], sd, gx.size), rng.normal(mean[1], sd, gx.size)], axis=1).astype(np.float64)
    vec[::5, 2:] = 0.0
    vec[3::11, :2] = -1.0
    vec[3::11, 2:] = 0.0
    return vec


@pytest.mark.parametrize("size,ps,channels", [((200, 150), 10, 3), ((320, 240), 10, 1), ((333, 211), 7, 3), ((1920, 1080), 10, 3), ((64, 48), 4, 1)])
def test_draw_flow_list_equals_oracle(capi, oracle, size, ps, channels):
    w, h = size
    rng = np.random.default_rng(w + ps)
    img = rng.integers(0, 256, (h, w, 3) if channels == 3 else (h, w), dtype=np.uint8)
    vec = _field(rng, w, h, ps, (5, -3), 0.35 * 5 * ps)          # long enough to overlap the neighbours and to leave the image
    col = (255, 0, 0) if channels == 3 else (180,)
    ctx = capi.Context(width=w, height=h, pixel_step=ps, min_vector_size=0.2)
    got, n = ctx.draw_flow(img, vec4=vec, colour=col)
    ref, m = oracle.draw_flow(img, vec, ps, 0.2, col)
    assert n == m and n > 20
    bad = (got != ref).any(axis=2) if channels == 3 else got != ref
    # the arrow-head end points go through atan2f / cos / sin (device libm vs glibc, <= 2 ulp): a head stroke may land one pixel off
    # where a coordinate sits within 1e-6 of a rounding boundary; everything else is integer arithmetic
    assert bad.sum() <= 12, int(bad.sum())
    if w <= 333:
        assert bad.sum() == 0
    ctx.close()


def test_draw_flow_empty_and_degenerate(capi, oracle):
    w, h = 96, 64
    img = np.random.default_rng(0).integers(0, 256, (h, w, 3), dtype=np.uint8)
    ctx = capi.Context(width=w, height=h, pixel_step=10, min_vector_size=1.0)
    got, n = ctx.draw_flow(img, vec4=np.zeros((0, 4)))
    assert n == 0 and np.array_equal(got, img)
    vec = np.array([[10, 10, 0.5, 0.9], [20, 20, 60.0, 1.0], [30, 30, 49.9, -49.9], [90, 60, 30.0, 20.0], [5, 5, -20.0, -20.0],
                    [50, 30, 1.5, 0.0], [50, 30, 0.0, 1.5], [-1, -1, 0, 0]], np.float64)
    got, n = ctx.draw_flow(img, vec4=vec, colour=(1, 2, 3))
    ref, m = oracle.draw_flow(img, vec, 10, 1.0, (1, 2, 3))
    assert n == m == 5
    assert np.array_equal(got, ref)
    ctx.close()


def test_draw_flow_from_batch_outputs(capi, oracle):
    """vec4 == NULL: the flow field of optical_flow_calculator.cpp:78-117 is formed on the device from one pair's next / status / keep."""
    w, h = 640, 480
    frames, _ = synth.sequence(w, h, 3, seed=12, blobs=3)
    ctx = capi.Context(width=w, height=h, max_batch=2, pixel_step=10, min_vector_size=0.2, seed=4)
    r = ctx.process_batch(frames)
    pts = oracle.grid_points(w, h, 10)
    rgb = np.repeat(frames[1][..., None], 3, axis=2)
    for b in range(2):
        got, n = ctx.draw_flow(rgb, next_pts=r["next"][b], status=r["status"][b], keep=r["keep"][b], colour=(255, 0, 0))
        vec = oracle.flow_field_row_major(pts, r["next"][b], r["status"][b], r["keep"][b])
        ref, m = oracle.draw_flow(rgb, vec, 10, 0.2, (255, 0, 0))
        assert n == m and n > 1000 and n <= int((r["keep"][b] != 0).sum())
        assert ((got != ref).any(axis=2)).sum() <= 4
    ctx.close()
