"""spfsplatv2_b200 -- B200-native (sm_100a) differentiable Gaussian-splatting renderer + 2-D RoPE
behind SPFSplatV2's ``DecoderSplattingCUDA`` / ``render_cuda`` / ``diff_gauss_pose`` / ``curope``
surfaces.  Host side: Python/PyTorch plumbing over a C-ABI library (include/spfsplat.h)."""
from __future__ import annotations

import sys

__all__ = ["install_shims", "build"]


def build(verbose: bool = False) -> str:
    from ._lib import build as _b
    return _b(verbose)


def install_shims() -> None:
    """Register the drop-in modules under the names the reference imports
    (``import diff_gauss_pose`` cuda_splatting.py:5; ``import curope`` curope2d.py:6-9)."""
    from . import curope as _curope
    from . import diff_gauss_pose as _dgp
    sys.modules["diff_gauss_pose"] = _dgp
    sys.modules["curope"] = _curope
