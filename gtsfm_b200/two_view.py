"""Batched two-view seam: the device half of `TwoViewEstimator.run_2view` for a shard of pairs.

The reference runs, per pair and per Dask task (gtsfm/two_view_estimator.py:350-481, 846-886): `normalize_coordinates`
(a Python loop of gtsam calls over every keypoint, gtsfm/utils/features.py:41-51), `cv2.findEssentialMat` +
`cv2.recoverPose` (gtsfm/frontend/verifier/ransac.py:74-81, gtsfm/utils/verification.py:83) and then hands the inliers
to triangulation / bundle adjustment on the CPU.  Here a whole shard of pairs is walked on one GPU with every keypoint
and match tensor resident in HBM: calibration, hypothesis generation, scoring, local optimisation, the inlier mask
and the cheirality vote are the kernels of csrc/ransac.cu (`b2_ransac_essential_dev`), pair p's verification runs on
its own stream under pair p+1's matching, and only what the CPU back half consumes leaves the device: the verified rows
of the match array, R, t and the inlier ratio - exactly the `VerifierBase.verify` tuple (verifier_base.py:67-90).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterable, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from .gtsfm_api import Rot3, Unit3
from .pipeline import DeviceFeatures, DeviceFrontEnd
from .verifier import DEFAULT_SEED

MIN_MATCHES_E = 6  # opencv_verifier_base.py:77-79
MATCH_BATCH = 8    # pairs per lock-step LightGlue batch (the library maximum)


@dataclass
class TwoViewResult:
    i2Ri1: Optional[Rot3]
    i2Ui1: Optional[Unit3]
    v_corr_idxs: np.ndarray  # (n, 2) rows of the putative match array that survived verification (host)
    inlier_ratio_est_model: float
    num_putative: int


def _failure(num_putative: int) -> TwoViewResult:  # verifier_base.py:60-64
    return TwoViewResult(None, None, np.zeros((0, 2), np.int64), 0.0, num_putative)


class B200TwoViewBatch:
    """match (optional) -> calibrate -> RANSAC-5pt -> recoverPose for a list of pairs, device-resident.

    `intrinsics[i]` = (f, u0, v0) of image i (Cal3Bundler without distortion: what GTSfM's deep front-end configs use).
    """

    def __init__(self, front_end: DeviceFrontEnd, estimation_threshold_px: float = 4.0, seed: int = DEFAULT_SEED):
        self.fe = front_end
        self.threshold_px = float(estimation_threshold_px)
        self.seed = int(seed)

    def run(self, features: Mapping[int, DeviceFeatures], pairs: Iterable[Tuple[int, int]], intrinsics: Mapping[int, Sequence[float]],
            putative: Optional[Mapping[Tuple[int, int], torch.Tensor]] = None) -> Dict[Tuple[int, int], TwoViewResult]:
        """`putative[(i1, i2)]`: (k, 2) int64 device tensor of match indices; matched here with LightGlue when absent."""
        out: Dict[Tuple[int, int], TwoViewResult] = {}
        pending = []  # (pair, matches, future) in flight on the verification stream
        pairs = list(pairs)
        for c0 in range(0, len(pairs), MATCH_BATCH):
            chunk = pairs[c0:c0 + MATCH_BATCH]
            todo = [pr for pr in chunk if putative is None or pr not in putative]
            matched = dict(zip(todo, self.fe.match_batch([(features[i1], features[i2]) for i1, i2 in todo])))
            for i1, i2 in chunk:
                a, b = features[i1], features[i2]
                m = putative[(i1, i2)] if (i1, i2) not in matched else matched[(i1, i2)][0]
                k = int(m.shape[0])
                if k < MIN_MATCHES_E:
                    out[(i1, i2)] = _failure(k)
                    continue
                # this chunk's verifications run on their own stream / thread under the next chunk's matching
                pending.append(((i1, i2), m, self.fe.verify_async(a, b, m, intrinsics[i1], intrinsics[i2], self.threshold_px, self.seed)))
            while len(pending) > MATCH_BATCH:
                self._collect(pending.pop(0), out)
        for p in pending:
            self._collect(p, out)
        return out

    @staticmethod
    def _collect(item, out) -> None:
        pair, m, fut = item
        E, R, t, n_inl, mask = fut.result()
        k = int(m.shape[0])
        if E is None:
            out[pair] = _failure(k)
            return
        keep = mask.bool()
        rows = m[keep].cpu().numpy().astype(np.int64)  # only the verified rows cross PCIe
        out[pair] = TwoViewResult(Rot3(R), Unit3(t), rows, float(n_inl) / float(k), k)
