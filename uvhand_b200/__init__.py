"""uvhand_b200 -- B200-native (sm_100a) multi-scale deformable attention, a drop-in for the
MSDeformAttn operator of On-JungWoan/UVHand (reference tree: models/ops).

Public surface mirrors the reference's:

    from uvhand_b200.functions import MSDeformAttnFunction      # models/ops/functions/ms_deform_attn_func.py:21-39
    from uvhand_b200.modules import MSDeformAttn                # models/ops/modules/ms_deform_attn.py:30-140
    import MultiScaleDeformableAttention as MSDA                # the extension module (vision.cpp:13-16)

The extension is built in-tree by ``python -m uvhand_b200.build`` (or ``__graft_entry__.build()``).
There is no CPU or PyTorch fallback: if the native extension is missing, importing the
function/module layer raises.
"""
from __future__ import annotations

import importlib
import sys

__all__ = ["load_extension", "MSDeformAttnFunction", "MSDeformAttn"]

_EXT_NAME = "MultiScaleDeformableAttention"


def load_extension():
    """Import the native extension and make it importable under the reference's top-level name."""
    mod = sys.modules.get(_EXT_NAME)
    if mod is not None and hasattr(mod, "ms_deform_attn_forward"):
        return mod
    import torch  # noqa: F401  (the extension links against libtorch)
    try:
        mod = importlib.import_module(f"{__name__}.{_EXT_NAME}")
    except ImportError as exc:  # fail loudly -- no fallback path exists
        raise ImportError(
            f"uvhand_b200: native extension {_EXT_NAME} is not built ({exc}). "
            "Run `python -m uvhand_b200.build` (needs nvcc with sm_100a support).") from exc
    sys.modules[_EXT_NAME] = mod
    return mod


def __getattr__(name):
    if name == "MSDeformAttnFunction":
        from .functions import MSDeformAttnFunction
        return MSDeformAttnFunction
    if name == "MSDeformAttn":
        from .modules import MSDeformAttn
        return MSDeformAttn
    raise AttributeError(name)
