"""TEST INFRASTRUCTURE — CPU restatement of the reference SuperPoint path (not product code; never shipped).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this module.  It restates, with plain torch-CPU fp32 ops, what these reference lines compute:

  * model:   thirdparty/SuperGluePretrainedNetwork/models/superpoint.py:47-92 (nms, borders, sampling), :145-202
  * wrapper: gtsfm/frontend/detector_descriptor/superpoint.py:63-93, gtsfm/utils/images.py:15-40,
             gtsfm/common/keypoints.py:89-127

Pinned by ``oracle/make_golden.py`` against the unmodified reference modules run in the build container with the
same seeded weights (bit-exact on keypoints / scores, <=1e-6 on descriptors); the reference's own tests hold no
golden keypoints for this path, so that run is the anchor.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

ENCODER = ["conv1a", "conv1b", "P", "conv2a", "conv2b", "P", "conv3a", "conv3b", "P", "conv4a", "conv4b"]


def rgb_to_gray_u8(rgb: np.ndarray) -> np.ndarray:
    """cv2.COLOR_RGB2GRAY in integer form (utils/images.py:36-38): (9798 R + 19235 G + 3735 B + 2^14) >> 15.

    (The often-quoted 4899/9617/1868 >> 14 variant is NOT what cv2 4.13 computes for 8-bit inputs; checked
    bit-exactly against cv2 in oracle/make_golden.py.)
    """
    if rgb.ndim == 2:
        return rgb
    r = rgb[..., 0].astype(np.int32)
    g = rgb[..., 1].astype(np.int32)
    b = rgb[..., 2].astype(np.int32)
    return ((9798 * r + 19235 * g + 3735 * b + (1 << 14)) >> 15).astype(np.uint8)


def _t(sd, name):
    return torch.from_numpy(np.ascontiguousarray(sd[name]))


def nms_mask_scores(scores: torch.Tensor, radius: int) -> torch.Tensor:
    """superpoint.py:47-62 — equality-based 3-round NMS with (2r+1)^2 max-pools, -inf padding."""
    k = 2 * radius + 1

    def pool(x):
        return F.max_pool2d(x, kernel_size=k, stride=1, padding=radius)

    keep = scores == pool(scores)
    for _ in range(2):
        near_kept = pool(keep.float()) > 0
        rest = torch.where(near_kept, torch.zeros_like(scores), scores)
        fresh = rest == pool(rest)
        keep = keep | (fresh & ~near_kept)
    return torch.where(keep, scores, torch.zeros_like(scores))


def superpoint_forward(
    gray01: np.ndarray,
    sd: Dict[str, np.ndarray],
    nms_radius: int = 4,
    keypoint_threshold: float = 0.005,
    border: int = 4,
    intermediates: Optional[dict] = None,
) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(H,W) float32 in [0,1] -> keypoints (N,2) f32 (x,y) row-major order, scores (N,), descriptors (N,256)."""
    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(gray01, dtype=np.float32))[None, None]
        for name in ENCODER:  # superpoint.py:148-158
            if name == "P":
                x = F.max_pool2d(x, 2, 2)
            else:
                x = F.relu(F.conv2d(x, _t(sd, name + ".weight"), _t(sd, name + ".bias"), padding=1))
                if intermediates is not None:
                    intermediates[name] = x[0].numpy().copy()
        # detector head, superpoint.py:161-167
        cpa = F.relu(F.conv2d(x, _t(sd, "convPa.weight"), _t(sd, "convPa.bias"), padding=1))
        logits = F.conv2d(cpa, _t(sd, "convPb.weight"), _t(sd, "convPb.bias"))
        prob = F.softmax(logits, 1)[:, :-1]
        _, _, hc, wc = prob.shape
        heat = prob.permute(0, 2, 3, 1).reshape(1, hc, wc, 8, 8).permute(0, 1, 3, 2, 4).reshape(1, hc * 8, wc * 8)
        if intermediates is not None:
            intermediates["logits"] = logits[0].numpy().copy()
            intermediates["heat"] = heat[0].numpy().copy()
        heat = nms_mask_scores(heat, nms_radius)[0]
        if intermediates is not None:
            intermediates["nms"] = heat.numpy().copy()
        # superpoint.py:170-178,187
        rc = torch.nonzero(heat > keypoint_threshold)
        sc = heat[rc[:, 0], rc[:, 1]]
        h8, w8 = hc * 8, wc * 8
        ok = (rc[:, 0] >= border) & (rc[:, 0] < h8 - border) & (rc[:, 1] >= border) & (rc[:, 1] < w8 - border)
        rc, sc = rc[ok], sc[ok]
        kp = torch.flip(rc, [1]).float()
        # descriptor head, superpoint.py:190-196 and sample_descriptors :80-92 with align_corners=True
        cda = F.relu(F.conv2d(x, _t(sd, "convDa.weight"), _t(sd, "convDa.bias"), padding=1))
        dense = F.normalize(F.conv2d(cda, _t(sd, "convDb.weight"), _t(sd, "convDb.bias")), p=2, dim=1)
        if intermediates is not None:
            intermediates["dense_desc"] = dense[0].numpy().copy()
        g = kp - 8 / 2 + 0.5
        g = g / torch.tensor([w8 - 8 / 2 - 0.5, h8 - 8 / 2 - 0.5])[None]
        g = g * 2 - 1
        samp = F.grid_sample(dense, g.view(1, 1, -1, 2), mode="bilinear", align_corners=True)
        desc = F.normalize(samp.reshape(1, 256, -1), p=2, dim=1)[0]
    return kp.numpy(), sc.numpy(), np.ascontiguousarray(desc.numpy().T)


def detect_and_describe(
    image: np.ndarray,
    sd: Dict[str, np.ndarray],
    max_keypoints: int = 5000,
    mask: Optional[np.ndarray] = None,
) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Wrapper semantics (gtsfm/.../superpoint.py:63-93): gray, /255, forward, mask filter, argpartition top-k."""
    gray = rgb_to_gray_u8(image).astype(np.float32) / 255.0
    kp, sc, desc = superpoint_forward(gray, sd)
    if mask is not None:  # keypoints.py:112-127
        r = np.round(kp).astype(int)
        valid = np.flatnonzero(mask[r[:, 1], r[:, 0]] == 1)
        kp, sc, desc = kp[valid], sc[valid], desc[valid]
    if max_keypoints < len(kp):  # keypoints.py:101-110
        sel = np.argpartition(-sc, max_keypoints)[:max_keypoints]
        kp, sc, desc = kp[sel], sc[sel], desc[sel]
    return kp, sc, desc
