"""The reference's decoder surface on top of the batched B200 rasterizer.

Mirrors (names, argument meaning, return shapes, error behaviour):
  * ``render_cuda``               /root/reference/src/model/decoder/cuda_splatting.py:45-144
  * ``render_cuda_orthographic``  /root/reference/src/model/decoder/cuda_splatting.py:146-255
  * ``DecoderSplattingCUDA``      /root/reference/src/model/decoder/decoder_splatting_cuda.py:23-78
  * ``DecoderOutput`` / ``Gaussians`` / ``get_decoder``  decoder/decoder.py:18-21, model/types.py:7-14,
    decoder/__init__.py:4-12

Differences in mechanism, not in results: all b*v views go through ONE launch sequence; the
1/near scaling of means/scales, the SH [P,3,K] -> [P,K,3] transpose and the v-fold ``repeat`` of
every Gaussian tensor (decoder_splatting_cuda.py:58-64) are index math inside the kernels; no
``.item()`` host syncs per view.
"""
from __future__ import annotations

from dataclasses import dataclass
from math import isqrt
from typing import Literal, Optional

import torch
from torch import Tensor, nn

from .camera import camera_setup_cuda, get_projection_matrix, orthographic_setup
from .rasterizer import RasterSettings, rasterize_batched, rasterize_batched_head

DepthRenderingMode = Literal["depth", "log", "disparity", "relative_disparity"]


@dataclass
class Gaussians:
    means: Tensor        # [b, g, 3]
    covariances: Tensor  # [b, g, 3, 3]  (not read by the decoder)
    rotations: Tensor    # [b, g, 4]
    scales: Tensor       # [b, g, 3]
    harmonics: Tensor    # [b, g, 3, d_sh]
    opacities: Tensor    # [b, g]


@dataclass
class DecoderOutput:
    color: Tensor            # [b, v, 3, h, w]
    depth: Optional[Tensor]  # [b, v, h, w]


def _render_views(view, proj, tanfov, scale, image_shape, background_color, means, sh, opacities, rotations,
                  scales, use_sh, enable_cov_grad, enable_sh_grad, views_per_scene):
    h, w = image_shape
    n = sh.shape[-1]
    degree = isqrt(n) - 1
    settings = RasterSettings(image_height=h, image_width=w, sh_degree=degree, scale_modifier=1.0,
                              views_per_scene=views_per_scene, sh_layout_ck=use_sh,
                              enable_cov_grad=enable_cov_grad, enable_sh_grad=enable_sh_grad)
    if use_sh:
        shs, colors = sh, None                     # encoder-native [S,P,3,K]; transposed by index math
    else:
        shs, colors = None, sh[..., 0]             # [S,P,3]
    color, depth, _alpha, _radii = rasterize_batched(
        settings, means, scales, rotations, opacities, shs, colors, view, proj, tanfov,
        background_color, scale)
    return color, depth


def render_cuda(extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor, image_shape: tuple,
                background_color: Tensor, gaussian_means: Tensor, gaussian_covariances: Tensor,
                gaussian_sh_coefficients: Tensor, gaussian_opacities: Tensor, gaussian_rotations: Tensor,
                gaussian_scales: Tensor, scale_invariant: bool = True, use_sh: bool = True,
                enable_cov_grad: bool = False, enable_sh_grad: bool = False, views_per_scene: int = 1):
    """Same contract as the reference's ``render_cuda``: returns (images [B,3,H,W], depths [B,1,H,W]).

    With ``views_per_scene = v > 1`` the Gaussian tensors carry B/v scenes and view i reads scene
    i // v (what DecoderSplattingCUDA needs, without materialising v copies)."""
    assert use_sh or gaussian_sh_coefficients.shape[-1] == 1
    view, proj, tanfov, scale = camera_setup_cuda(extrinsics, intrinsics, near, far, scale_invariant)
    return _render_views(view, proj, tanfov, scale if scale_invariant else None, image_shape, background_color,
                         gaussian_means, gaussian_sh_coefficients, gaussian_opacities, gaussian_rotations,
                         gaussian_scales, use_sh, enable_cov_grad, enable_sh_grad, views_per_scene)


def render_cuda_orthographic(extrinsics: Tensor, width: Tensor, height: Tensor, near: Tensor, far: Tensor,
                             image_shape: tuple, background_color: Tensor, gaussian_means: Tensor,
                             gaussian_covariances: Tensor, gaussian_sh_coefficients: Tensor,
                             gaussian_opacities: Tensor, gaussian_rotations: Tensor, gaussian_scales: Tensor,
                             fov_degrees: float = 0.1, use_sh: bool = True, dump: Optional[dict] = None,
                             enable_cov_grad: bool = False, enable_sh_grad: bool = False) -> Tensor:
    """Fake-orthographic render (tiny fov, camera moved back); returns images only, like the reference."""
    assert use_sh or gaussian_sh_coefficients.shape[-1] == 1
    view, proj, tanfov, info = orthographic_setup(extrinsics, width, height, near, far, fov_degrees)
    if dump is not None:
        dump.update(info)
    color, _ = _render_views(view, proj, tanfov, None, image_shape, background_color, gaussian_means,
                             gaussian_sh_coefficients, gaussian_opacities, gaussian_rotations, gaussian_scales,
                             use_sh, enable_cov_grad, enable_sh_grad, 1)
    return color


@dataclass
class DecoderSplattingCUDACfg:
    name: Literal["splatting_cuda"]
    background_color: list
    make_scale_invariant: bool
    enable_cov_grad: bool
    enable_sh_grad: bool


class DecoderSplattingCUDA(nn.Module):
    def __init__(self, cfg: DecoderSplattingCUDACfg) -> None:
        super().__init__()
        self.cfg = cfg
        self.make_scale_invariant = cfg.make_scale_invariant
        self.enable_cov_grad = cfg.enable_cov_grad
        self.enable_sh_grad = cfg.enable_sh_grad
        self.register_buffer("background_color", torch.tensor(cfg.background_color, dtype=torch.float32),
                             persistent=False)

    def forward(self, gaussians: Gaussians, extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
                image_shape: tuple, depth_mode: Optional[str] = None) -> DecoderOutput:
        b, v = extrinsics.shape[:2]
        h, w = image_shape
        color, depth = render_cuda(
            extrinsics.reshape(b * v, 4, 4), intrinsics.reshape(b * v, 3, 3), near.reshape(b * v),
            far.reshape(b * v), image_shape, self.background_color.expand(b * v, 3),
            gaussians.means, gaussians.covariances, gaussians.harmonics, gaussians.opacities,
            gaussians.rotations, gaussians.scales, scale_invariant=self.make_scale_invariant,
            enable_cov_grad=self.enable_cov_grad, enable_sh_grad=self.enable_sh_grad, views_per_scene=v)
        color = color.view(b, v, 3, h, w)
        depth = depth.view(b, v, h, w)
        if self.make_scale_invariant:
            depth = depth * near[:, :, None, None]
        return DecoderOutput(color, depth)

    def forward_head(self, means: Tensor, head_out: Tensor, extrinsics: Tensor, intrinsics: Tensor, near: Tensor,
                     far: Tensor, image_shape: tuple, sh_degree: int, opacity_exponent: float = 1.0,
                     eps: float = 1e-8) -> DecoderOutput:
        """Encoder head rows straight to images: ``head_out`` [b, g, 1 + 7 + 3*d_sh] is the Gaussian
        head's output (density logit in channel 0), ``means`` [b, g, 3].  Equivalent to
        ``self.forward(adapter.forward_head(means, head_out, ...), ...)`` -- encoder_spfsplatv2.py:255-268 followed by
        decoder_splatting_cuda.py:40-78 -- but the adapter's and the opacity mapping's outputs (and their gradients) never
        touch HBM: 2 x 332 bytes per Gaussian less traffic each way, two kernels fewer per step."""
        b, v = extrinsics.shape[:2]
        h, w = image_shape
        view, proj, tanfov, scale = camera_setup_cuda(extrinsics.reshape(b * v, 4, 4), intrinsics.reshape(b * v, 3, 3),
                                                      near.reshape(b * v), far.reshape(b * v), self.make_scale_invariant)
        settings = RasterSettings(image_height=h, image_width=w, sh_degree=sh_degree, views_per_scene=v,
                                  enable_cov_grad=self.enable_cov_grad, enable_sh_grad=self.enable_sh_grad)
        color, depth, _a, _r = rasterize_batched_head(settings, means, head_out, view, proj, tanfov,
                                                      self.background_color.expand(b * v, 3),
                                                      scale if self.make_scale_invariant else None,
                                                      eps=eps, opacity_exponent=opacity_exponent)
        depth = depth.view(b, v, h, w)
        if self.make_scale_invariant:
            depth = depth * near[:, :, None, None]
        return DecoderOutput(color.view(b, v, 3, h, w), depth)


DECODERS = {"splatting_cuda": DecoderSplattingCUDA}
DecoderCfg = DecoderSplattingCUDACfg


def get_decoder(decoder_cfg: DecoderCfg) -> DecoderSplattingCUDA:
    return DECODERS[decoder_cfg.name](decoder_cfg)
