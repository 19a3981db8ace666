"""Import the *unmodified* reference modules from /root/reference (authoring
container only -- the tree does not exist on the GPU box).

TEST INFRASTRUCTURE.  Used by ``oracle/make_goldens.py`` and by the optional
``tests/test_oracle_vs_reference.py`` (skipped when the tree is absent).  Nothing
here copies reference source; it only arranges ``sys.path`` / ``sys.modules`` so
that the reference's own files import on this machine:

  1. ``timm`` is not installed: a stand-in ``timm.models.layers`` exporting
     ``to_2tuple``, ``trunc_normal_`` and an identity ``DropPath`` (swin_512.py:4).
  2. ``contrast.resnet`` does not exist in the tree (contrast/option.py:3): empty stub.
  3. On CPU ``Tensor.cuda`` becomes a no-op because posMask / negMask /
     regression_loss hard-code ``.cuda()`` (PixPro_swin_v5.py:54-55,65-67,89-104).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("STSWIN_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "seg18/net/Ours/swin_512.py"))


def _install_timm_standin() -> None:
    if "timm" in sys.modules:
        return
    import torch.nn as nn

    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")

    class DropPath(nn.Identity):
        def __init__(self, drop_prob=None):
            super().__init__()

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    layers.DropPath = DropPath
    layers.to_2tuple = to_2tuple
    layers.trunc_normal_ = nn.init.trunc_normal_
    timm.models = models
    models.layers = layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})


def _cuda_noop_on_cpu() -> None:
    import torch
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self


def import_swin():
    """-> module seg18/net/Ours/swin_512.py"""
    assert available(), f"reference tree not found at {REF_ROOT}"
    _install_timm_standin()
    path = os.path.join(REF_ROOT, "seg18")
    if path not in sys.path:
        sys.path.insert(0, path)
    return importlib.import_module("net.Ours.swin_512")


def import_pixpro():
    """-> module pixcontrast_18/contrast/models/PixPro_swin_v5.py"""
    assert available(), f"reference tree not found at {REF_ROOT}"
    _install_timm_standin()
    _cuda_noop_on_cpu()
    path = os.path.join(REF_ROOT, "pixcontrast_18")
    if path not in sys.path:
        sys.path.insert(0, path)
    if "contrast.resnet" not in sys.modules:
        import contrast  # the reference package itself
        stub = types.ModuleType("contrast.resnet")
        stub.__all__ = []
        sys.modules["contrast.resnet"] = stub
        contrast.resnet = stub
    return importlib.import_module("contrast.models.PixPro_swin_v5")
