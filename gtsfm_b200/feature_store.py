"""Compact feature cache / wire format.

The reference caches `(Keypoints, descriptors)` per image as a bz2-compressed pickle (gtsfm/frontend/cacher/
detector_descriptor_cacher.py:71-95 -> gtsfm/utils/io.py write_to_bz2_file): 5 MB of float32 through single-threaded bz2
(~1 s per image, ~7 % saved), and the same pickles travel between Dask workers.  `pack_features` / `unpack_features` define a
flat little-endian record instead - fixed header, coordinates as uint16 when they are integral (SuperPoint's always are:
keypoints are pixel centres), responses float32, descriptors float32 (lossless, the default) or float16 (opt-in, half the
bytes, |error| <= 2^-12 on unit-norm descriptors: NOT bit-identical matches) - that can be written, mmapped or sent as is.
`B200DetectorDescriptorCacher` is the drop-in for the reference cacher around any DetectorDescriptorBase: same constructor,
same cache key (class name + the reference's image hash), same `detect_and_describe` contract."""
from __future__ import annotations

import hashlib
import struct
from pathlib import Path
from typing import Optional, Tuple

import numpy as np

from .gtsfm_api import DetectorDescriptorBase, Image, Keypoints

MAGIC = b"B2F1"
_FLAG_U16_COORDS, _FLAG_F16_DESC, _FLAG_SCALES, _FLAG_RESPONSES = 1, 2, 4, 8
_HEADER = struct.Struct("<4sIIII")  # magic, N, D, flags, reserved


def pack_features(keypoints: Keypoints, descriptors: np.ndarray, desc_dtype: str = "f32") -> bytes:
    """-> one self-describing byte string.  Lossless for desc_dtype == "f32"."""
    coords = np.ascontiguousarray(keypoints.coordinates, np.float32).reshape(-1, 2)
    n = coords.shape[0]
    desc = np.ascontiguousarray(descriptors, np.float32).reshape(n, -1) if n else np.zeros((0, 0), np.float32)
    flags = 0
    integral = n > 0 and np.array_equal(coords, np.rint(coords)) and coords.min() >= 0 and coords.max() < 65536
    if integral:
        flags |= _FLAG_U16_COORDS
    if desc_dtype == "f16":
        flags |= _FLAG_F16_DESC
    elif desc_dtype != "f32":
        raise ValueError("desc_dtype must be 'f32' or 'f16'")
    parts = []
    if keypoints.scales is not None:
        flags |= _FLAG_SCALES
    if keypoints.responses is not None:
        flags |= _FLAG_RESPONSES
    parts.append(_HEADER.pack(MAGIC, n, desc.shape[1] if n else 0, flags, 0))
    parts.append((coords.astype("<u2") if integral else coords.astype("<f4")).tobytes())
    if keypoints.scales is not None:
        parts.append(np.ascontiguousarray(keypoints.scales, "<f4").tobytes())
    if keypoints.responses is not None:
        parts.append(np.ascontiguousarray(keypoints.responses, "<f4").tobytes())
    parts.append((desc.astype("<f2") if desc_dtype == "f16" else desc.astype("<f4")).tobytes())
    return b"".join(parts)


def unpack_features(buf: bytes) -> Tuple[Keypoints, np.ndarray]:
    magic, n, d, flags, _ = _HEADER.unpack_from(buf, 0)
    if magic != MAGIC:
        raise ValueError("not a B2F1 feature record")
    off = _HEADER.size

    def take(dtype, count, shape):
        nonlocal off
        a = np.frombuffer(buf, dtype=dtype, count=count, offset=off).reshape(shape)
        off += a.nbytes
        return a

    coords = take("<u2" if flags & _FLAG_U16_COORDS else "<f4", 2 * n, (n, 2)).astype(np.float32)
    scales = take("<f4", n, (n,)).copy() if flags & _FLAG_SCALES else None
    responses = take("<f4", n, (n,)).copy() if flags & _FLAG_RESPONSES else None
    desc = take("<f2" if flags & _FLAG_F16_DESC else "<f4", n * d, (n, d)).astype(np.float32)
    return Keypoints(coords, scales=scales, responses=responses), desc


def image_hash(image: Image) -> str:
    """gtsfm/utils/cache.py:14-23 generate_hash_for_image: sha1(file name, width, height) + sha1(pixel bytes)."""
    arr = np.ascontiguousarray(image.value_array)
    meta = "{}_{}_{}".format(getattr(image, "file_name", None), arr.shape[1], arr.shape[0]).encode()
    return hashlib.sha1(meta).hexdigest() + hashlib.sha1(arr.tobytes()).hexdigest()


class B200DetectorDescriptorCacher(DetectorDescriptorBase):
    """Drop-in for gtsfm/frontend/cacher/detector_descriptor_cacher.py:28-100 writing B2F1 records instead of bz2 pickles."""

    def __init__(self, detector_descriptor_obj: DetectorDescriptorBase, cache_root: Optional[Path] = None, desc_dtype: str = "f32") -> None:
        super().__init__(max_keypoints=detector_descriptor_obj.max_keypoints)
        self._detector_descriptor = detector_descriptor_obj
        self._detector_descriptor_obj_cache_key = type(detector_descriptor_obj).__name__
        self._root = Path(cache_root) if cache_root is not None else Path.cwd() / "cache"
        self._desc_dtype = desc_dtype

    def _path(self, image: Image) -> Path:
        return self._root / "detector_descriptor" / f"{self._detector_descriptor_obj_cache_key}_{image_hash(image)}.b2f"

    def detect_and_describe(self, image: Image) -> Tuple[Keypoints, np.ndarray]:
        path = self._path(image)
        if path.exists():
            return unpack_features(path.read_bytes())
        keypoints, descriptors = self._detector_descriptor.detect_and_describe(image)
        path.parent.mkdir(parents=True, exist_ok=True)
        tmp = path.with_suffix(".tmp")
        tmp.write_bytes(pack_features(keypoints, descriptors, self._desc_dtype))
        tmp.replace(path)  # atomic: concurrent workers never read a half-written record
        return keypoints, descriptors
