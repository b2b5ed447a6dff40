"""GPU: the batched two-view seam (gtsfm_b200/two_view.py) returns, pair by pair, what the per-pair plugins return."""
import numpy as np
import pytest
import torch

from gtsfm_b200 import synthetic as syn
from gtsfm_b200.gtsfm_api import Cal3Bundler, Keypoints
from gtsfm_b200.pipeline import DeviceFrontEnd
from gtsfm_b200.two_view import B200TwoViewBatch
from gtsfm_b200.verifier import B200Ransac

pytestmark = pytest.mark.gpu


def test_batch_equals_per_pair_plugins():
    fe = DeviceFrontEnd(syn.superpoint_state_dict(0), syn.lightglue_state_dict(2, "sharp"), max_keypoints=1024)
    frames, cal = syn.synthetic_sequence(4, 240, 320)
    feats = {i: fe.detect(torch.from_numpy(f).cuda()) for i, f in enumerate(frames)}
    pairs = [(0, 1), (0, 2), (1, 3), (2, 3)]
    intr = {i: cal for i in feats}
    res = B200TwoViewBatch(fe, 4.0).run(feats, pairs, intr)
    assert set(res) == set(pairs)
    ver = B200Ransac(True, 4.0)
    calib = Cal3Bundler(cal[0], 0, 0, cal[1], cal[2])
    n_ok = 0
    for (i1, i2), r in res.items():
        m, _ = fe.match(feats[i1], feats[i2])
        m = m.cpu().numpy()
        assert r.num_putative == len(m)
        k1, k2 = Keypoints(feats[i1].kp.cpu().numpy()), Keypoints(feats[i2].kp.cpu().numpy())
        R, U, v, ratio = ver.verify(k1, k2, m, calib, calib)
        if R is None:
            assert r.i2Ri1 is None and len(r.v_corr_idxs) == 0
            continue
        n_ok += 1
        assert np.array_equal(r.v_corr_idxs, v), "verified rows differ from the per-pair plugin"
        np.testing.assert_allclose(r.i2Ri1.matrix(), R.matrix(), atol=1e-9)
        np.testing.assert_allclose(r.i2Ui1.point3(), U.point3(), atol=1e-9)
        assert abs(r.inlier_ratio_est_model - ratio) < 1e-12
    assert n_ok >= 2, "the synthetic sequence should give verifiable pairs"


def test_too_few_matches_is_the_failure_tuple():
    fe = DeviceFrontEnd(syn.superpoint_state_dict(0), syn.lightglue_state_dict(2, "sharp"), max_keypoints=256)
    frames, cal = syn.synthetic_sequence(2, 120, 160)
    feats = {i: fe.detect(torch.from_numpy(f).cuda()) for i, f in enumerate(frames)}
    put = {(0, 1): torch.zeros((3, 2), dtype=torch.int64, device="cuda")}
    r = B200TwoViewBatch(fe).run(feats, [(0, 1)], {0: cal, 1: cal}, putative=put)[(0, 1)]
    assert r.i2Ri1 is None and r.i2Ui1 is None and r.v_corr_idxs.shape == (0, 2) and r.num_putative == 3
