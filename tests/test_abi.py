"""CPU suite: the C-ABI library builds, loads without a GPU, and exports every symbol include/gtsfm_b200.h declares."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from gtsfm_b200 import build

    path = build.build()
    return ctypes.CDLL(str(path))


def _declared_symbols():
    text = (ROOT / "include" / "gtsfm_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/gtsfm_b200.h but not exported: {missing}"


def test_python_binding_covers_header(lib):
    from gtsfm_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_version_and_no_gpu_failure_is_loud(lib):
    import torch

    lib.b2_version.restype = ctypes.c_int
    assert lib.b2_version() >= 100
    if not torch.cuda.is_available():
        from gtsfm_b200 import _lib

        with pytest.raises(_lib.B200Error):
            _lib.Context(0)


def test_weight_packing_shapes():
    from gtsfm_b200 import synthetic as syn
    from gtsfm_b200 import weights

    assert weights.pack_superpoint(syn.superpoint_state_dict(0)).size == 1300865
    lg = weights.pack_lightglue(syn.lightglue_state_dict(2))
    assert lg.dtype == np.float32 and lg.size == 11851601
    sg = weights.pack_superglue(syn.superglue_state_dict(1))
    assert sg.dtype == np.float32 and sg.size == 12003905
