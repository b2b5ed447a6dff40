"""B200 LightGlue / SuperGlue matcher plugins.

Drop-ins for gtsfm/frontend/matcher/lightglue_matcher.py:24-112 (`LightGlueMatcher`) and
gtsfm/frontend/matcher/superglue_matcher.py:30-115 (`SuperGlueMatcher`): same `match(...)` signature, argument
meaning, errors and output dtypes ((K, 2) int64 for LightGlue, uint32 for SuperGlue, rows ascending in column 0).
All arithmetic runs in libgtsfm_b200.so; lazily created device state keeps the objects picklable
(tests/frontend/matcher/test_matcher_base.py:102-107).
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional, Tuple, Union

import numpy as np

from . import _lib, weights
from .gtsfm_api import Keypoints, MatcherBase

DESC_DIM = 256


class LightGlueEngine:
    def __init__(self, state_dict, device: int = 0, ctx: Optional[_lib.Context] = None, feature_cache: Optional[bool] = None):
        self.ctx = ctx or _lib.Context(device)
        if feature_cache is not None:
            self.ctx.set_option("feature_cache", 1 if feature_cache else 0)
        blob = weights.pack_lightglue(weights.load_state_dict(state_dict))
        self.ctx.check(self.ctx.lib.b2_lightglue_set_weights(self.ctx.handle, _lib.ptr(blob), blob.size), "lightglue_set_weights")
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def match(self, kp0, desc0, kp1, desc1, depth_confidence=0.95, width_confidence=0.99, filter_threshold=0.1,
              prune_min_kpts=-1, return_scores=False, fp16_attention=False):
        kp0 = np.ascontiguousarray(kp0, np.float32)
        kp1 = np.ascontiguousarray(kp1, np.float32)
        desc0 = np.ascontiguousarray(desc0, np.float32)
        desc1 = np.ascontiguousarray(desc1, np.float32)
        n0, n1 = len(kp0), len(kp1)
        cap = max(1, min(n0, n1))
        out = np.empty((cap, 2), np.int64)
        sc = np.empty(cap, np.float32)
        k, stop = _lib.C.c_int(0), _lib.C.c_int(0)
        prm = _lib.LightGlueParams(depth_confidence, width_confidence, filter_threshold, prune_min_kpts, 1 if fp16_attention else 0)
        sent0 = self.ctx.h2d_bytes()
        rc = self.ctx.lib.b2_lightglue_match_host(self.ctx.handle, _lib.ptr(kp0), _lib.ptr(desc0), n0, _lib.ptr(kp1),
                                                  _lib.ptr(desc1), n1, _lib.C.byref(prm), _lib.ptr(out), _lib.ptr(sc),
                                                  _lib.C.byref(k), _lib.C.byref(stop))
        self.ctx.check(rc, "lightglue_match")
        self.last_stop = stop.value
        self.h2d_bytes += self.ctx.h2d_bytes() - sent0  # feature arrays the library already holds are not re-sent
        self.d2h_bytes += k.value * (16 + 4) + 8
        if return_scores:
            return out[: k.value].copy(), sc[: k.value].copy()
        return out[: k.value].copy()


class B200LightGlueMatcher(MatcherBase):
    """LightGlue on hand-written sm_100a kernels behind GTSfM's MatcherBase.

    `cpu_semantics=True` (default) reproduces what the reference computes on its CPU front-end (pruning attempted at
    every layer, lightglue.py:339-344) and is what the parity fixtures pin; False uses the reference's CUDA+flash
    pruning threshold of 1536 keypoints.

    `feature_cache=True` (opt-in, default off) lets the library keep device copies of the host feature arrays it is handed,
    keyed by (host address, size) and validated by a hash of the full contents: an image matched against many partners is
    uploaded once.  It only helps callers that pass the SAME numpy buffers repeatedly (a sequential in-process loop);
    under Dask every task unpickles fresh arrays and the default - copy on every call, like the reference - is what runs.
    """

    def __init__(self, features: str = "superpoint", use_cuda: bool = True, weights_path: Union[Path, str, dict, None] = None,
                 device: int = 0, cpu_semantics: bool = True, feature_cache: bool = False, fp16_attention: bool = False):
        super().__init__()
        if features != "superpoint":
            raise ValueError(f"Unsupported features: {features} (this build serves the SuperPoint LightGlue only)")
        if weights_path is None:
            raise FileNotFoundError("LightGlue weights_path is required (superpoint_lightglue_v0-1_arxiv.pth-style checkpoint)")
        if not isinstance(weights_path, dict) and not Path(weights_path).exists():
            raise FileNotFoundError(f"LightGlue weights not found at {weights_path}")
        self._use_cuda = use_cuda
        self._features = features
        self._weights = weights_path
        self._device = device
        self._cpu_semantics = cpu_semantics
        self._feature_cache = bool(feature_cache)
        self._fp16_attention = bool(fp16_attention)  # opt-in: the reference's CUDA numerics (fp16 flash SDPA), ~2x faster attention
        self._engine: Optional[LightGlueEngine] = None

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_engine"] = None
        return st

    def _ensure_engine(self) -> LightGlueEngine:
        if self._engine is None:
            self._engine = LightGlueEngine(self._weights, self._device, feature_cache=self._feature_cache)
        return self._engine

    def match(self, keypoints_i1: Keypoints, keypoints_i2: Keypoints, descriptors_i1: np.ndarray, descriptors_i2: np.ndarray,
              im_shape_i1: Tuple[int, int, int], im_shape_i2: Tuple[int, int, int]) -> np.ndarray:
        if keypoints_i1.responses is None or keypoints_i2.responses is None:
            raise ValueError("Responses for keypoints required for LightGlue.")  # lightglue_matcher.py:78-79
        if len(keypoints_i1) == 0 or len(keypoints_i2) == 0:
            return np.zeros((0, 2), np.int64)
        if descriptors_i1.shape[1] != DESC_DIM or descriptors_i2.shape[1] != DESC_DIM:
            raise AssertionError("LightGlue(superpoint) expects 256-dimensional descriptors")  # lightglue.py:509-510
        eng = self._ensure_engine()
        return eng.match(keypoints_i1.coordinates, descriptors_i1, keypoints_i2.coordinates, descriptors_i2,
                         prune_min_kpts=-1 if self._cpu_semantics else 1536, fp16_attention=self._fp16_attention)


SUPERGLUE_DESC_DIM = 256
DEFAULT_NUM_SINKHORN_ITERATIONS = 20  # superglue_matcher.py:27
SUPERGLUE_MATCH_THRESHOLD = 0.2  # thirdparty/.../superglue.py:201


class SuperGlueEngine:
    def __init__(self, state_dict, device: int = 0, ctx: Optional[_lib.Context] = None):
        self.ctx = ctx or _lib.Context(device)
        blob = weights.pack_superglue(weights.load_state_dict(state_dict))
        self.ctx.check(self.ctx.lib.b2_superglue_set_weights(self.ctx.handle, _lib.ptr(blob), blob.size), "superglue_set_weights")
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def match(self, kp0, sc0, desc0, kp1, sc1, desc1, shape0, shape1, sinkhorn_iters=DEFAULT_NUM_SINKHORN_ITERATIONS,
              match_threshold=SUPERGLUE_MATCH_THRESHOLD, return_scores=False):
        arrs = [np.ascontiguousarray(a, np.float32) for a in (kp0, sc0, desc0, kp1, sc1, desc1)]
        n0, n1 = len(arrs[0]), len(arrs[3])
        cap = max(1, min(n0, n1))
        out = np.empty((cap, 2), np.uint32)
        sc = np.empty(cap, np.float32)
        k = _lib.C.c_int(0)
        rc = self.ctx.lib.b2_superglue_match_host(self.ctx.handle, _lib.ptr(arrs[0]), _lib.ptr(arrs[1]), _lib.ptr(arrs[2]), n0, int(shape0[0]),
                                                  int(shape0[1]), _lib.ptr(arrs[3]), _lib.ptr(arrs[4]), _lib.ptr(arrs[5]), n1, int(shape1[0]),
                                                  int(shape1[1]), int(sinkhorn_iters), float(match_threshold), _lib.ptr(out), _lib.ptr(sc),
                                                  _lib.C.byref(k))
        self.ctx.check(rc, "superglue_match")
        self.h2d_bytes += sum(a.nbytes for a in arrs)
        self.d2h_bytes += k.value * 12 + 4
        if return_scores:
            return out[: k.value].copy(), sc[: k.value].copy()
        return out[: k.value].copy()


class B200SuperGlueMatcher(MatcherBase):
    """SuperGlue on hand-written sm_100a kernels behind GTSfM's MatcherBase (gtsfm/frontend/matcher/superglue_matcher.py:30-115)."""

    def __init__(self, use_cuda: bool = True, use_outdoor_model: bool = True, weights_path: Union[Path, str, dict, None] = None, device: int = 0):
        super().__init__()
        if weights_path is None:
            raise FileNotFoundError("SuperGlue weights_path is required (a superglue_outdoor.pth-style checkpoint)")
        if not isinstance(weights_path, dict) and not Path(weights_path).exists():
            raise FileNotFoundError(f"SuperGlue weights not found at {weights_path}")
        self._use_cuda = use_cuda
        self._config = {"descriptor_dim": SUPERGLUE_DESC_DIM, "weights": "outdoor" if use_outdoor_model else "indoor",
                        "sinkhorn_iterations": DEFAULT_NUM_SINKHORN_ITERATIONS}
        self._weights = weights_path
        self._device = device
        self._engine: Optional[SuperGlueEngine] = None

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_engine"] = None
        return st

    def _ensure_engine(self) -> SuperGlueEngine:
        if self._engine is None:
            self._engine = SuperGlueEngine(self._weights, self._device)
        return self._engine

    def match(self, keypoints_i1: Keypoints, keypoints_i2: Keypoints, descriptors_i1: np.ndarray, descriptors_i2: np.ndarray,
              im_shape_i1: Tuple[int, int, int], im_shape_i2: Tuple[int, int, int]) -> np.ndarray:
        if keypoints_i1.responses is None or keypoints_i2.responses is None:
            raise ValueError("Responses for keypoints required for SuperGlue")  # superglue_matcher.py:78-79
        if len(keypoints_i1) == 0 or len(keypoints_i2) == 0:
            return np.zeros((0, 2), np.uint32)
        if descriptors_i1.shape[1] != SUPERGLUE_DESC_DIM or descriptors_i2.shape[1] != SUPERGLUE_DESC_DIM:
            raise Exception("Superglue pretrained network only works on 256 dimensional descriptors")  # :81-82
        eng = self._ensure_engine()
        return eng.match(keypoints_i1.coordinates, keypoints_i1.responses, descriptors_i1, keypoints_i2.coordinates,
                         keypoints_i2.responses, descriptors_i2, im_shape_i1, im_shape_i2, self._config["sinkhorn_iterations"])
