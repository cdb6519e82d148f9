"""Drop-in module for the external ``diff_gauss_pose`` package the reference imports at
/root/reference/src/model/decoder/cuda_splatting.py:5.

Same public names and call shapes as recorded at the reference's call sites
(cuda_splatting.py:105-138, 218-249):

    settings = GaussianRasterizationSettings(image_height=..., image_width=..., tanfovx=..., tanfovy=...,
                   bg=..., scale_modifier=..., projmatrix=..., sh_degree=..., prefiltered=..., debug=...,
                   enable_cov_grad=..., enable_sh_grad=...)
    image, depth, norm, alpha, radii, extra = GaussianRasterizer(settings)(
                   means3D=..., means2D=..., shs=... | colors_precomp=..., opacities=..., scales=...,
                   rotations=..., viewmatrix=...)

Put ``spfsplatv2_b200`` *as a search path* on sys.path (or call ``spfsplatv2_b200.install_shims()``)
and the unmodified reference decoder runs on libspfsplat.so.  One view per call here (that is the
reference's contract); the batched entry is ``spfsplatv2_b200.decoder.render_cuda``.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch
from torch import Tensor, nn

from ..rasterizer import RasterSettings, rasterize_batched


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: Tensor
    scale_modifier: float
    projmatrix: Tensor
    sh_degree: int
    prefiltered: bool
    debug: bool
    enable_cov_grad: bool = True
    enable_sh_grad: bool = True


def _scalar(v) -> float:
    # cuda_splatting.py:108-109 passes python floats, :221-222 passes 0-dim / 1-element tensors
    if isinstance(v, Tensor):
        return float(v.reshape(-1)[0])
    return float(v)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D: Tensor, means2D: Optional[Tensor] = None, opacities: Tensor = None,
                shs: Optional[Tensor] = None, colors_precomp: Optional[Tensor] = None,
                scales: Optional[Tensor] = None, rotations: Optional[Tensor] = None,
                cov3Ds_precomp: Optional[Tensor] = None, viewmatrix: Tensor = None, extra_attrs=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if cov3Ds_precomp is not None:
            raise NotImplementedError("cov3Ds_precomp is not supported (the reference never passes it, "
                                      "cuda_splatting.py:136)")
        if scales is None or rotations is None:
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        if extra_attrs is not None:
            raise NotImplementedError("extra_attrs is not supported")
        if viewmatrix is None:
            raise Exception("viewmatrix is required")
        dev = means3D.device
        tanfov = torch.tensor([[_scalar(rs.tanfovx), _scalar(rs.tanfovy)]], dtype=torch.float32, device=dev)
        settings = RasterSettings(
            image_height=int(rs.image_height), image_width=int(rs.image_width), sh_degree=int(rs.sh_degree),
            scale_modifier=float(rs.scale_modifier), views_per_scene=1, sh_layout_ck=False,
            enable_cov_grad=bool(rs.enable_cov_grad), enable_sh_grad=bool(rs.enable_sh_grad),
            want_alpha=True, want_means2d_grad=means2D is not None and means2D.requires_grad)
        color, depth, alpha, radii = rasterize_batched(
            settings, means3D[None], scales[None], rotations[None], opacities.reshape(1, -1),
            None if shs is None else shs[None], None if colors_precomp is None else colors_precomp[None],
            viewmatrix[None], rs.projmatrix[None], tanfov, rs.bg.reshape(1, 3), None,
            None if means2D is None else means2D[None])
        return color[0], depth[0], None, alpha[0], radii[0], None
