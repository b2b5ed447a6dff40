"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference modules from /root/reference (this container only).

Used by ``oracle/make_golden.py`` to pin the CPU restatement in ``oracle/*_ref.py`` and to write the committed
fixtures under ``tests/golden``.  Never imported by the product (``gtsfm_b200``), by ``-m gpu`` tests, ``smoke()`` or
``bench.py``: /root/reference does not exist on the GPU box.

Shims:
  * the thirdparty models read checkpoints from fixed paths / URLs (superpoint.py:136-137, superglue.py:221-224,
    lightglue.py:415-421): ``torch.load`` / ``torch.hub.load_state_dict_from_url`` are patched for the duration of
    construction to return the seeded synthetic state dict;
  * ``lightglue.py`` is loaded by file path (its package ``__init__`` needs kornia);
  * ``sample_descriptors`` picks ``align_corners`` from ``int(torch.__version__[2]) > 2`` (superpoint.py:87), which
    misreads "2.11.0" as 1; the pinned torch 2.7.0 yields True, so the reference run is forced to True by presenting
    a version string "2.7.0" to that module only.
"""
from __future__ import annotations

import contextlib
import importlib.util
import sys
from pathlib import Path
from unittest import mock

import numpy as np
import torch

REF = Path("/root/reference")
SGPN = REF / "thirdparty" / "SuperGluePretrainedNetwork"
LG = REF / "thirdparty" / "LightGlue" / "lightglue" / "lightglue.py"


def available() -> bool:
    return SGPN.exists() and LG.exists()


def _to_torch(sd):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


def _load_by_path(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, str(path))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def ref_superpoint(state_dict, **config):
    mod = _load_by_path("_ref_superpoint", SGPN / "models" / "superpoint.py")
    with mock.patch.object(torch, "load", lambda *a, **k: _to_torch(state_dict)), contextlib.redirect_stdout(None):
        model = mod.SuperPoint(config).eval()

    # force the pinned-torch behaviour of superpoint.py:87 (align_corners=True): the module sees a torch whose
    # version string is the pinned 2.7.0; every other attribute falls through to the real torch.
    class _Proxy:
        __version__ = "2.7.0"

        def __getattr__(self, item):
            return getattr(torch, item)

    mod.torch = _Proxy()
    return model


def ref_superglue(state_dict, **config):
    mod = _load_by_path("_ref_superglue", SGPN / "models" / "superglue.py")
    with mock.patch.object(torch, "load", lambda *a, **k: _to_torch(state_dict)), contextlib.redirect_stdout(None):
        model = mod.SuperGlue(config).eval()
    return model


def ref_lightglue(state_dict, **conf):
    mod = _load_by_path("_ref_lightglue", LG)
    with mock.patch.object(torch.hub, "load_state_dict_from_url", lambda *a, **k: _to_torch(state_dict)):
        model = mod.LightGlue(features="superpoint", **conf).eval()
    return model
