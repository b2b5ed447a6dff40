"""GPU parity: CUDA SuperPoint (through the C ABI / plugin) vs the golden fixtures written by the reference and vs the
CPU oracle on seeded inputs.  Bars (BASELINE.json north_star): keypoint indices exact, descriptors within 1e-3."""
import pickle

import numpy as np
import pytest

from gtsfm_b200 import synthetic as syn
from gtsfm_b200.detector_descriptor import B200SuperPointDetectorDescriptor, SuperPointEngine
from gtsfm_b200.gtsfm_api import Image

pytestmark = pytest.mark.gpu

DESC_TOL = 1e-3  # north_star: "within 1e-3 on descriptors"
SCORE_TOL = 1e-5


@pytest.fixture(scope="module")
def engine(b200_ctx):
    return SuperPointEngine(syn.superpoint_state_dict(0), ctx=b200_ctx)


def _gray(name, golden_dir):
    from oracle.superpoint_ref import rgb_to_gray_u8

    fx = np.load(golden_dir / f"superpoint_{name}.npz")
    if "gray" in fx:
        return fx["gray"], fx
    frames = {"tiny": (0, 120, 160), "odd": (3, 203, 317), "vga": (1, 480, 640), "mp1": (2, 1024, 1024)}
    idx, h, w = frames[name]
    return rgb_to_gray_u8(syn.synthetic_frame(idx, h, w)), fx


@pytest.mark.parametrize("name", ["tiny", "odd", "vga", "mp1", "lund1", "lund2"])  # mp1 = 1024 x 1024 (BASELINE configs[2])
def test_detect_matches_reference_fixture(engine, golden_dir, name):
    gray, fx = _gray(name, golden_dir)
    xy, sc = engine.detect(gray)
    ref_xy = fx["keypoints"].astype(np.float32)
    assert len(xy) == len(ref_xy), f"{name}: {len(xy)} keypoints vs reference {len(ref_xy)}"
    assert np.array_equal(xy, ref_xy), f"{name}: keypoint coordinates / order differ"
    np.testing.assert_allclose(sc, fx["scores"], rtol=0, atol=SCORE_TOL)
    rows = fx["desc_rows"]
    desc = engine.describe(xy[rows])
    assert np.abs(desc - fx["desc"]).max() < DESC_TOL
    full = engine.describe(xy)
    assert abs(float(full.astype(np.float64).sum()) - float(fx["desc_checksum"])) < DESC_TOL * full.size * 1e-2
    np.testing.assert_allclose(np.linalg.norm(full, axis=1), 1.0, atol=1e-5)


def test_intermediate_maps_match_oracle(engine, b200_ctx, golden_dir):
    from oracle.superpoint_ref import superpoint_forward

    gray, _ = _gray("odd", golden_dir)
    inter = {}
    superpoint_forward(gray.astype(np.float32) / 255.0, syn.superpoint_state_dict(0), intermediates=inter)
    engine.detect(gray)
    hc, wc = gray.shape[0] // 8, gray.shape[1] // 8
    heat = b200_ctx.debug_fetch("heat", hc * wc * 64).reshape(hc * 8, wc * 8)
    np.testing.assert_allclose(heat, inter["heat"], atol=1e-6)
    nms = b200_ctx.debug_fetch("nms", hc * wc * 64).reshape(hc * 8, wc * 8)
    assert np.array_equal(nms > 0, inter["nms"] > 0), "NMS survivor set differs"
    feat = b200_ctx.debug_fetch("conv4b", hc * wc * 128).reshape(hc, wc, 128).transpose(2, 0, 1)
    np.testing.assert_allclose(feat, inter["conv4b"], atol=2e-5, rtol=1e-5)
    dense = b200_ctx.debug_fetch("dense_desc", hc * wc * 256).reshape(hc, wc, 256).transpose(2, 0, 1)
    np.testing.assert_allclose(dense, inter["dense_desc"], atol=1e-5)


def test_rgb_input_and_plugin_contract(golden_dir, tmp_path):
    """Reference contract tests (tests/frontend/detector/test_detector_base.py:27-56,
    tests/frontend/detector_descriptor/test_detector_descriptor_base.py:29-42) + wrapper parity incl. top-k order."""
    from oracle.superpoint_ref import detect_and_describe

    sd = syn.superpoint_state_dict(0)
    wpath = tmp_path / "superpoint_v1.pth"
    syn.save_pth(sd, wpath)
    det = B200SuperPointDetectorDescriptor(max_keypoints=500, weights_path=wpath)
    pickle.dumps(det)
    rgb = syn.synthetic_frame(7, 240, 320)
    kps, desc = det.detect_and_describe(Image(rgb))
    pickle.dumps(det)  # still picklable after the engine exists
    assert len(kps) <= 500 and len(kps) == desc.shape[0] and desc.shape[1] == 256
    assert np.all(kps.coordinates[:, 0] >= 0) and np.all(kps.coordinates[:, 0] <= 320)
    assert np.all(kps.coordinates[:, 1] >= 0) and np.all(kps.coordinates[:, 1] <= 240)
    okp, osc, odesc = detect_and_describe(rgb, sd, 500)
    assert np.array_equal(kps.coordinates, okp), "top-k selection / order differs from the reference wrapper"
    np.testing.assert_allclose(kps.responses, osc, atol=SCORE_TOL)
    assert np.abs(desc - odesc).max() < DESC_TOL
    # mask filter (gtsfm/.../superpoint.py:87-89)
    mask = np.zeros((240, 320), np.uint8)
    mask[:, :160] = 1
    kps_m, desc_m = det.detect_and_describe(Image(rgb, mask=mask))
    okp_m, _, _ = detect_and_describe(rgb, sd, 500, mask=mask)
    assert np.array_equal(kps_m.coordinates, okp_m) and len(desc_m) == len(okp_m)
    # run-to-run exactness (tests/repro_tests/.../test_detector_descriptor_reproducibility_base.py:31-36)
    for _ in range(10):
        k2, d2 = det.detect_and_describe(Image(rgb))
        assert k2 == kps and np.array_equal(d2, desc)


def test_missing_weights_raise(tmp_path):
    with pytest.raises(FileNotFoundError):
        B200SuperPointDetectorDescriptor(weights_path=tmp_path / "nope.pth")


def test_stale_map_token_is_refused(b200_ctx, golden_dir):
    """Two images interleaved on one handle: describe with the first image's token after a second detect must fail loudly
    instead of sampling the wrong dense map."""
    from gtsfm_b200._lib import B200Error

    eng = SuperPointEngine(syn.superpoint_state_dict(0), ctx=b200_ctx)
    gray_a, _ = _gray("tiny", golden_dir)
    gray_b, _ = _gray("odd", golden_dir)
    xy_a, _ = eng.detect(gray_a)
    tok_a = eng.map_token
    d_a = eng.describe(xy_a[:16], tok_a)
    eng.detect(gray_b)
    with pytest.raises(B200Error, match="stale feature-map token"):
        eng.describe(xy_a[:16], tok_a)
    eng.detect(gray_a)
    np.testing.assert_array_equal(eng.describe(xy_a[:16]), d_a)


def test_exact_fp32_simt_path(golden_dir):
    """SIMT fp32 convolutions (set_option force_simt) reproduce the reference keypoints as well."""
    from gtsfm_b200 import _lib

    ctx = _lib.Context(0)
    try:
        ctx.set_option("force_simt", 1)
        eng = SuperPointEngine(syn.superpoint_state_dict(0), ctx=ctx)
        gray, fx = _gray("odd", golden_dir)
        xy, sc = eng.detect(gray)
        assert np.array_equal(xy, fx["keypoints"].astype(np.float32))
        assert np.abs(eng.describe(xy[fx["desc_rows"]]) - fx["desc"]).max() < DESC_TOL
    finally:
        ctx.close()
