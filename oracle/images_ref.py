"""TEST INFRASTRUCTURE - CPU restatement of the loader's image ingest arithmetic.

The reference resizes every image with `cv2.resize(..., interpolation=cv2.INTER_CUBIC)` to the size chosen by
`get_downsampling_factor_per_axis` (gtsfm/utils/images.py:102-129,150-220, called from gtsfm/loader/loader_base.py:160-200) and
later converts to gray with cv2's fixed-point RGB2GRAY (gtsfm/utils/images.py:15-40).  OpenCV (opencv-python, pinned 4.12.0.88,
not vendored under /root/reference) implements the uint8 cubic resize in 11-bit fixed point; this file restates that published
arithmetic in numpy and `tests/test_images_cpu.py` pins it bit-for-bit against the installed cv2."""
from __future__ import annotations

from typing import Tuple

import numpy as np

COEF_BITS = 11  # INTER_RESIZE_COEF_BITS
COEF_SCALE = 1 << COEF_BITS


def downsampled_size(img_h: int, img_w: int, max_resolution: int) -> Tuple[int, int]:
    """gtsfm/utils/images.py:150-220: the short side becomes max_resolution when it is larger; else unchanged."""
    if min(img_h, img_w) <= max_resolution:
        return img_h, img_w
    if img_h <= img_w:
        return max_resolution, int(np.round(img_w * (max_resolution / float(img_h))).astype(np.int32))
    return int(np.round(img_h * (max_resolution / float(img_w))).astype(np.int32)), max_resolution


def cubic_taps(n_dst: int, n_src: int) -> Tuple[np.ndarray, np.ndarray]:
    """Per destination index: first source index (may be < 0 or > n_src - 4: clamped when read) and the four int16 weights.
    cv::resize, INTER_CUBIC: fx = (dx + 0.5) * scale - 0.5 in float32, A = -0.75, weights rounded to 11-bit fixed point."""
    scale = float(n_src) / float(n_dst)
    dx = np.arange(n_dst, dtype=np.float64)
    fx = ((dx + 0.5) * scale - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int64)
    fx = (fx - sx.astype(np.float32)).astype(np.float32)
    A = np.float32(-0.75)
    one, two = np.float32(1), np.float32(2)
    c0 = ((A * (fx + one) - np.float32(5) * A) * (fx + one) + np.float32(8) * A) * (fx + one) - np.float32(4) * A
    c1 = ((A + two) * fx - (A + np.float32(3))) * fx * fx + one
    c2 = ((A + two) * (one - fx) - (A + np.float32(3))) * (one - fx) * (one - fx) + one
    c3 = one - c0 - c1 - c2
    w = np.stack([c0, c1, c2, c3], -1).astype(np.float32) * np.float32(COEF_SCALE)
    wi = np.clip(np.rint(w), -32768, 32767).astype(np.int16)  # saturate_cast<short>(float): round half to even
    return sx - 1, wi


def resize_cubic_u8(img: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_CUBIC) for uint8 HxW or HxWxC."""
    src = img if img.ndim == 3 else img[:, :, None]
    h, w, _ = src.shape
    x0, wx = cubic_taps(new_w, w)
    y0, wy = cubic_taps(new_h, h)
    s = src.astype(np.int64)
    hor = np.zeros((h, new_w, src.shape[2]), np.int64)
    for k in range(4):
        hor += s[:, np.clip(x0 + k, 0, w - 1), :] * wx[:, k].astype(np.int64)[None, :, None]
    acc = np.zeros((new_h, new_w, src.shape[2]), np.int64)
    for k in range(4):
        acc += hor[np.clip(y0 + k, 0, h - 1), :, :] * wy[:, k].astype(np.int64)[:, None, None]
    out = np.clip((acc + (1 << (2 * COEF_BITS - 1))) >> (2 * COEF_BITS), 0, 255).astype(np.uint8)
    return out if img.ndim == 3 else out[:, :, 0]
