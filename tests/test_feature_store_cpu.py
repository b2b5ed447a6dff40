"""CPU: the compact feature record (gtsfm_b200/feature_store.py) round-trips losslessly and the cacher honours the
DetectorDescriptorBase contract (reference cacher: gtsfm/frontend/cacher/detector_descriptor_cacher.py:71-95)."""
import bz2
import pickle

import numpy as np

from gtsfm_b200 import synthetic as syn
from gtsfm_b200.feature_store import B200DetectorDescriptorCacher, pack_features, unpack_features
from gtsfm_b200.gtsfm_api import DetectorDescriptorBase, Image, Keypoints


def _features(n=5000):
    kp0, sc0, d0, *_ = syn.synthetic_features(3, n, 8)
    return Keypoints(np.rint(kp0).astype(np.float32), responses=sc0), d0


def test_round_trip_is_lossless_and_smaller_than_bz2_pickle():
    kps, desc = _features()
    rec = pack_features(kps, desc)
    k2, d2 = unpack_features(rec)
    assert np.array_equal(k2.coordinates, kps.coordinates) and k2.coordinates.dtype == np.float32
    assert np.array_equal(k2.responses, kps.responses) and k2.scales is None
    assert np.array_equal(d2, desc) and d2.dtype == np.float32
    ref = bz2.compress(pickle.dumps({"keypoints": kps, "descriptors": desc}))  # gtsfm/utils/io.py write_to_bz2_file
    assert len(rec) < 1.1 * len(ref), (len(rec), len(ref))
    half = pack_features(kps, desc, "f16")
    k3, d3 = unpack_features(half)
    assert len(half) < 0.55 * len(rec) and np.abs(d3 - desc).max() <= 2.0 ** -11
    # non-integral coordinates and empty sets survive too
    odd = Keypoints(kps.coordinates + 0.25, scales=np.ones(len(kps), np.float32))
    k4, _ = unpack_features(pack_features(odd, desc))
    assert np.array_equal(k4.coordinates, odd.coordinates) and np.array_equal(k4.scales, odd.scales) and k4.responses is None
    k5, d5 = unpack_features(pack_features(Keypoints(np.zeros((0, 2), np.float32)), np.zeros((0, 256), np.float32)))
    assert len(k5) == 0 and d5.size == 0


class _Counting(DetectorDescriptorBase):
    def __init__(self):
        super().__init__(max_keypoints=100)
        self.calls = 0

    def detect_and_describe(self, image):
        self.calls += 1
        kps, desc = _features(100)
        return kps, desc


def test_cacher_hits_and_misses(tmp_path):
    inner = _Counting()
    cacher = B200DetectorDescriptorCacher(inner, cache_root=tmp_path)
    img_a, img_b = Image(syn.synthetic_frame(0, 60, 80)), Image(syn.synthetic_frame(1, 60, 80))
    a1 = cacher.detect_and_describe(img_a)
    a2 = cacher.detect_and_describe(img_a)
    cacher.detect_and_describe(img_b)
    assert inner.calls == 2 and cacher.max_keypoints == 100
    assert a1[0] == a2[0] and np.array_equal(a1[1], a2[1])
    assert len(list((tmp_path / "detector_descriptor").glob("_Counting_*.b2f"))) == 2
    pickle.dumps(cacher)
