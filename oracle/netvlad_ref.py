"""TEST INFRASTRUCTURE (CPU oracle; never imported by the product path).

Torch-CPU restatement of the reference's NetVLAD global descriptor, thirdparty/hloc/netvlad.py:
  * `forward`       :163-193  image * 255, clamp, - mean; VGG16 features[:-2] (13 convolutions, ReLU after all but the last, 4 max-
                               pools; :106-109); per-location L2 normalisation; NetVLAD layer; whitening Linear + final L2 normalisation
  * `netvlad_layer` :52-75    soft assignment = softmax over K = 64 clusters of a 1x1 projection, sum of assignment-weighted
                               residuals to the centres, intra-normalisation per cluster, flatten (d major, k minor), L2 normalise
Pinned by tests/golden/netvlad.npz, made by oracle/make_golden.py from the reference module's own forward (seeded weights).
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from gtsfm_b200.synthetic import NETVLAD_CONVS, NETVLAD_POOL_AFTER


def netvlad_layer(x: torch.Tensor, score_w: torch.Tensor, centers: torch.Tensor) -> torch.Tensor:
    """x: (B, 512, N) unit-norm columns -> (B, 32768)."""
    b = x.size(0)
    scores = F.softmax(F.conv1d(x, score_w), dim=1)  # (B, 64, N)
    diff = x.unsqueeze(2) - centers.unsqueeze(0).unsqueeze(-1)  # (B, 512, 64, N)
    desc = (scores.unsqueeze(1) * diff).sum(dim=-1)  # (B, 512, 64)
    desc = F.normalize(desc, dim=1)
    return F.normalize(desc.view(b, -1), dim=1)


def netvlad_forward(sd: Dict[str, np.ndarray], images: np.ndarray) -> np.ndarray:
    """images: (B, 3, H, W) float32 in [0, 1] -> (B, 4096) global descriptors."""
    t = {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
    with torch.no_grad():
        x = torch.clamp(torch.from_numpy(np.asarray(images, np.float32)) * 255, 0.0, 255.0)
        x = x - t["mean"].view(1, -1, 1, 1)
        last = NETVLAD_CONVS[-1][0]
        for idx, _, _ in NETVLAD_CONVS:
            x = F.conv2d(x, t[f"backbone.{idx}.weight"], t[f"backbone.{idx}.bias"], padding=1)
            if idx != last:
                x = F.relu(x)
            if idx in NETVLAD_POOL_AFTER:
                x = F.max_pool2d(x, 2, 2)
        b, c = x.shape[:2]
        x = F.normalize(x.view(b, c, -1), dim=1)
        d = netvlad_layer(x, t["netvlad.score_proj.weight"], t["netvlad.centers"])
        d = F.normalize(F.linear(d, t["whiten.weight"], t["whiten.bias"]), dim=1)
    return d.numpy()
