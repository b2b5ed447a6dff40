"""TEST INFRASTRUCTURE — CPU restatement of the reference verifier path (never shipped).

The reference's verifier arithmetic lives in a third-party dependency that is not under /root/reference:
**opencv-python, pinned 4.12.0.88 (uv.lock:1911-1912); this image has cv2 4.13.0** (recorded with every fixture).
This module therefore drives ``cv2`` exactly the way the reference does

  * gtsfm/frontend/verifier/opencv_verifier_base.py:46-110 (guards, fx = max, inlier rows, ratio)
  * gtsfm/frontend/verifier/ransac.py:74-81 (findEssentialMat USAC_ACCURATE, thr/fx, prob 0.999999) and :103-110
  * gtsfm/utils/features.py:41-51 (Cal3Bundler.calibrate with k1 = k2 = 0)
  * gtsfm/utils/verification.py:54-96 (recoverPose), :99-112 (E = K2^T F K1)

and restates in numpy the published quantities the CUDA verifier is judged on (squared Sampson distance, pose
angles) plus the scene of tests/frontend/verifier/test_verifier_base.py:235-294.  USAC's sampler / local
optimisation are not reproducible bit-for-bit (SURVEY.md §7 hard part 4), so parity for this row is the reference
tests' own criteria: pose within 2 deg, all rows returned on noise-free data, failure tuple on degenerate input,
plus inlier-set agreement with cv2 on seeded noisy scenes.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def calibrate(coords: np.ndarray, fx: float, u0: float, v0: float) -> np.ndarray:
    """Cal3Bundler.calibrate for k1 = k2 = 0 (utils/features.py:51): ((u - u0) / f, (v - v0) / f), float64."""
    c = np.asarray(coords, np.float64)
    return np.stack([(c[:, 0] - u0) / fx, (c[:, 1] - v0) / fx], -1)


def sampson_sq(E: np.ndarray, x1: np.ndarray, x2: np.ndarray) -> np.ndarray:
    """Squared Sampson distance of x2^T E x1 = 0 for (K,2) points."""
    p1 = np.concatenate([x1, np.ones((len(x1), 1))], 1)
    p2 = np.concatenate([x2, np.ones((len(x2), 1))], 1)
    l2 = p1 @ E.T
    l1 = p2 @ E
    num = np.sum(p2 * l2, 1) ** 2
    return num / (l2[:, 0] ** 2 + l2[:, 1] ** 2 + l1[:, 0] ** 2 + l1[:, 1] ** 2)


def rot_angle_deg(Ra: np.ndarray, Rb: np.ndarray) -> float:
    c = (np.trace(Ra.T @ Rb) - 1) / 2
    return float(np.degrees(np.arccos(np.clip(c, -1, 1))))


def dir_angle_deg(ta: np.ndarray, tb: np.ndarray) -> float:
    ta, tb = ta / np.linalg.norm(ta), tb / np.linalg.norm(tb)
    return float(np.degrees(np.arccos(np.clip(ta @ tb, -1, 1))))


def verify_cv2(
    kp1: np.ndarray, kp2: np.ndarray, matches: np.ndarray, K1: Tuple[float, float, float], K2: Tuple[float, float, float],
    use_intrinsics: bool = True, thr_px: float = 4.0,
) -> Tuple[Optional[np.ndarray], Optional[np.ndarray], np.ndarray, float, Optional[np.ndarray]]:
    """-> (R (3,3) | None, t (3,) | None, inlier rows of ``matches``, inlier ratio, E).  K = (f, u0, v0)."""
    import cv2

    fail = (None, None, np.array([], dtype=np.uint64), 0.0, None)
    if matches.shape[0] < (5 if use_intrinsics else 8):
        return fail
    if use_intrinsics:
        n1, n2 = calibrate(kp1, *K1), calibrate(kp2, *K2)
        if matches.shape[0] < 6:
            return fail
        fx = max(K1[0], K2[0])
        E, mask = cv2.findEssentialMat(n1[matches[:, 0]], n2[matches[:, 1]], np.eye(3), method=cv2.USAC_ACCURATE,
                                       threshold=thr_px / fx, prob=0.999999)
    else:
        Fm, mask = cv2.findFundamentalMat(np.asarray(kp1)[matches[:, 0]], np.asarray(kp2)[matches[:, 1]],
                                          method=cv2.FM_RANSAC, ransacReprojThreshold=thr_px, confidence=0.999999,
                                          maxIters=1000000)
        Km = lambda k: np.array([[k[0], 0, k[1]], [0, k[0], k[2]], [0, 0, 1.0]])
        E = Km(K2).T @ Fm @ Km(K1)
    inl = np.where(mask.ravel() == 1)[0]
    rows = matches[inl]
    ratio = float(np.mean(mask))
    _, R, t, _ = cv2.recoverPose(E, calibrate(kp1[rows[:, 0]], *K1), calibrate(kp2[rows[:, 1]], *K2))
    return R, t.ravel(), rows, ratio, E


def pose_to_euler_zyx_and_i1ti2(i2Ri1: np.ndarray, i2ti1: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """The quantities the reference's Argoverse known-answer test compares (tests/frontend/verifier/
    test_verifier_argoverse.py:86-104): invert the relative pose, Euler angles 'zyx' in degrees of i1Ri2, and i1ti2."""
    from scipy.spatial.transform import Rotation

    i1Ri2 = np.asarray(i2Ri1, np.float64).T
    t = np.asarray(i2ti1, np.float64).ravel()
    i1ti2 = -i1Ri2 @ (t / np.linalg.norm(t))
    return Rotation.from_matrix(i1Ri2).as_euler("zyx", degrees=True), i1ti2


def _Rx(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def _Ry(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def two_planes_scene(m: int, n: int, seed: int = 15):
    """tests/frontend/verifier/test_verifier_base.py:235-294 restated without gtsam.

    -> (uv1 (m+n,2), uv2 (m+n,2), i2Ri1 (3,3), i2ti1 unit (3,)), cameras with f = 1, u0 = v0 = 0.
    """
    rng = np.random.default_rng(seed)

    def on_plane(co, k):
        a, b, c, d = co
        x = rng.uniform(-5, 7, k)
        y = rng.uniform(-10, 10, k)
        return np.stack([x, y, -(a * x + b * y + d) / c], -1)

    pts = np.vstack([on_plane((-10, -1, -20, 150), m), on_plane((15, -2, -35, 200), n)])
    wt1, wt2 = np.array([0.1, 0, -20.0]), np.array([1, -2, -20.4])
    wR1, wR2 = _Rx(np.pi / 20), _Ry(np.pi / 6)
    c1 = (pts - wt1) @ wR1  # rows: wR^T (p - t)
    c2 = (pts - wt2) @ wR2
    uv1, uv2 = c1[:, :2] / c1[:, 2:], c2[:, :2] / c2[:, 2:]
    R21 = wR2.T @ wR1
    t21 = wR2.T @ (wt1 - wt2)
    return uv1, uv2, R21, t21 / np.linalg.norm(t21)


def synthetic_two_view(seed: int, k: int, inlier_ratio: float, noise_px: float = 0.5, f: float = 800.0,
                       width: int = 1280, height: int = 960):
    """Seeded two-view scene (SURVEY.md §8d verifier micro-benchmarks).

    -> (kp1 (k,2) px, kp2 (k,2) px, matches (k,2) uint32 identity rows, K=(f,u0,v0), R21, t21 unit, is_inlier (k,))
    """
    rng = np.random.default_rng(seed)
    u0, v0 = width / 2, height / 2
    ang = rng.uniform(-0.25, 0.25, 3)
    R21 = _Rx(ang[0]) @ _Ry(ang[1]) @ np.array([[np.cos(ang[2]), -np.sin(ang[2]), 0], [np.sin(ang[2]), np.cos(ang[2]), 0], [0, 0, 1]])
    t21 = rng.normal(0, 1, 3)
    t21[2] *= 0.3
    t21 /= np.linalg.norm(t21)
    n_in = int(round(k * inlier_ratio))
    # points in front of camera 1, depth 4..12 baselines
    x1 = np.stack([rng.uniform(-0.7, 0.7, k), rng.uniform(-0.5, 0.5, k)], -1)
    depth = rng.uniform(4, 12, k)
    P1 = np.concatenate([x1, np.ones((k, 1))], 1) * depth[:, None]
    P2 = P1 @ R21.T + t21
    x2 = P2[:, :2] / P2[:, 2:]
    kp1 = x1 * f + [u0, v0]
    kp2 = x2 * f + [u0, v0]
    kp1 = kp1 + rng.normal(0, noise_px, kp1.shape)
    kp2 = kp2 + rng.normal(0, noise_px, kp2.shape)
    is_in = np.zeros(k, bool)
    is_in[:n_in] = True
    kp2[n_in:] = np.stack([rng.uniform(0, width, k - n_in), rng.uniform(0, height, k - n_in)], -1)
    perm = rng.permutation(k)
    kp1, kp2, is_in = kp1[perm], kp2[perm], is_in[perm]
    matches = np.stack([np.arange(k), np.arange(k)], -1).astype(np.uint32)
    return kp1.astype(np.float64), kp2.astype(np.float64), matches, (f, u0, v0), R21, t21, is_in
