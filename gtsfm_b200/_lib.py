"""ctypes binding of libgtsfm_b200.so (include/gtsfm_b200.h).  No fallback: a missing library or GPU raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "libgtsfm_b200.so"

_lib: Optional[C.CDLL] = None


class B200Error(RuntimeError):
    pass


class LightGlueParams(C.Structure):
    _fields_ = [("depth_confidence", C.c_double), ("width_confidence", C.c_double), ("filter_threshold", C.c_double),
                ("prune_min_kpts", C.c_int), ("fp16_attention", C.c_int)]


class LightGluePair(C.Structure):
    _fields_ = [("kp0", C.c_void_p), ("desc0", C.c_void_p), ("n0", C.c_int), ("kp1", C.c_void_p), ("desc1", C.c_void_p), ("n1", C.c_int),
                ("out_matches", C.c_void_p), ("out_scores", C.c_void_p), ("out_k", C.c_int), ("out_stop_layer", C.c_int)]


class RansacParams(C.Structure):
    _fields_ = [("threshold", C.c_double), ("confidence", C.c_double), ("max_iters", C.c_int), ("seed", C.c_uint64)]


_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
_ip = C.POINTER(C.c_int)

# name -> (restype, argtypes): every symbol include/gtsfm_b200.h declares
SIGNATURES = {
    "b2_version": (_i, []),
    "b2_create": (_i, [_i, C.POINTER(_vp)]),
    "b2_destroy": (None, [_vp]),
    "b2_last_error": (C.c_char_p, [_vp]),
    "b2_launch_count": (C.c_uint64, [_vp]),
    "b2_h2d_bytes": (C.c_uint64, [_vp]),
    "b2_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    "b2_profile_start": (_i, [_vp, C.c_char_p]),
    "b2_profile_stop": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
    "b2_debug_fetch": (C.c_int64, [_vp, C.c_char_p, _vp, C.c_int64]),
    "b2_debug_gemm_host": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i]),
    "b2_superpoint_set_weights": (_i, [_vp, _vp, _sz]),
    "b2_superpoint_detect_dev": (_i, [_vp, _vp, _i, _i, _i, _sz, _f, _i, _i, _vp, _vp, _i, _ip, C.POINTER(C.c_uint64), _vp]),
    "b2_superpoint_describe_dev": (_i, [_vp, C.c_uint64, _vp, _i, _vp, _vp]),
    "b2_superpoint_extract_dev": (_i, [_vp, _vp, _i, _i, _i, _sz, _f, _i, _i, _i, _vp, _vp, _vp, _ip, _vp]),
    "b2_superpoint_extract_async_dev": (_i, [_vp, _vp, _i, _i, _i, _sz, _f, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "b2_superpoint_finish_dev": (_i, [_vp, _vp]),
    "b2_topk_indices_dev": (_i, [_vp, _vp, _i, _i, _vp, _ip, _vp]),
    "b2_superpoint_detect_host": (_i, [_vp, _vp, _i, _i, _i, _f, _i, _i, _vp, _vp, _i, _ip, C.POINTER(C.c_uint64)]),
    "b2_superpoint_describe_host": (_i, [_vp, C.c_uint64, _vp, _i, _vp]),
    "b2_image_resize_dev": (_i, [_vp, _vp, _i, _i, _i, _sz, _vp, _i, _i, _vp]),
    "b2_lightglue_set_weights": (_i, [_vp, _vp, _sz]),
    "b2_lightglue_match_dev": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, C.POINTER(LightGlueParams), _vp, _vp, _ip, _ip, _vp]),
    "b2_lightglue_match_host": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, C.POINTER(LightGlueParams), _vp, _vp, _ip, _ip]),
    "b2_lightglue_match_batched_dev": (_i, [_vp, C.POINTER(LightGluePair), _i, C.POINTER(LightGlueParams), _vp]),
    "b2_superglue_set_weights": (_i, [_vp, _vp, _sz]),
    "b2_superglue_match_dev": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _ip, _vp]),
    "b2_superglue_match_host": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _ip]),
    "b2_netvlad_blob_floats": (_sz, []),
    "b2_netvlad_set_weights": (_i, [_vp, _vp, _sz]),
    "b2_netvlad_describe_dev": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "b2_netvlad_describe_host": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "b2_similarity_pairs_host": (_i, [_vp, _vp, _i, _i, _i, _f, _vp, _vp]),
    "b2_ransac_essential_host": (_i, [_vp, _vp, _vp, _i, C.POINTER(RansacParams), _vp, _vp, _ip, _vp, _vp]),
    "b2_ransac_fundamental_host": (_i, [_vp, _vp, _vp, _i, C.POINTER(RansacParams), _vp, _vp, _ip]),
    "b2_ransac_essential_dev": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, C.POINTER(RansacParams), _vp, _vp, _ip, _vp, _vp, _vp]),
    "b2_recover_pose_host": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _ip]),
}


def load() -> C.CDLL:
    """dlopen the in-tree library and bind every declared symbol.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise B200Error(f"{LIB_PATH} is missing: run `python -m gtsfm_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a) -> C.c_void_p:
    """Raw pointer of a C-contiguous numpy array, a torch tensor (data_ptr) or an int address."""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous(), "tensor must be contiguous"
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(int(a))


class Context:
    """Owns one b2_context (one CUDA device, one internal stream).  Created lazily by the plugins."""

    def __init__(self, device: int = 0):
        self._lib = load()
        h = C.c_void_p()
        rc = self._lib.b2_create(int(device), C.byref(h))
        if rc != 0 or not h:
            raise B200Error(
                f"b2_create(device={device}) failed with {rc}: an sm_100 (B200) GPU is required; there is no CPU fallback")
        self.handle = h
        self.device = device

    def check(self, rc: int, what: str) -> None:
        if rc < 0:
            raise B200Error(f"{what} failed ({rc}): {self._lib.b2_last_error(self.handle).decode()}")

    @property
    def lib(self) -> C.CDLL:
        return self._lib

    def launch_count(self) -> int:
        return int(self._lib.b2_launch_count(self.handle))

    def set_option(self, name: str, value: int) -> None:
        self.check(self._lib.b2_set_option(self.handle, name.encode(), int(value)), f"set_option({name})")

    def h2d_bytes(self) -> int:
        return int(self._lib.b2_h2d_bytes(self.handle))

    def profile_start(self, kernel_prefix: str) -> None:
        self.check(self._lib.b2_profile_start(self.handle, kernel_prefix.encode()), "profile_start")

    def profile_stop(self):
        ms, n, w = C.c_double(0), C.c_uint64(0), C.c_double(0)
        self.check(self._lib.b2_profile_stop(self.handle, C.byref(ms), C.byref(n), C.byref(w)), "profile_stop")
        return ms.value, int(n.value), w.value

    def debug_fetch(self, name: str, max_floats: int) -> np.ndarray:
        out = np.empty(max_floats, np.float32)
        n = self._lib.b2_debug_fetch(self.handle, name.encode(), ptr(out), max_floats)
        self.check(int(n), f"debug_fetch({name})")
        return out[:n]

    def close(self) -> None:
        if getattr(self, "handle", None):
            self._lib.b2_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
