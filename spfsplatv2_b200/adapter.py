"""Fused ``UnifiedGaussianAdapter`` over ``spf_adapter_forward/backward``.

Mirrors /root/reference/src/model/encoder/common/gaussian_adapter.py:122-150: raw head output ``[..., 7 + 3*d_sh]`` plus
means and opacities -> ``Gaussians``.  One kernel each way instead of ~10 elementwise torch kernels; the ``[..., 3, 3]``
covariances, which the splatting decoder never reads (cuda_splatting.py:136 is commented out), are returned as a
stride-0 zero view instead of being computed and stored (9 floats per Gaussian).  CUDA tensors only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch
from torch import Tensor

from . import _lib as L
from .decoder import Gaussians


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class _Adapter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw: Tensor, d_sh: int, eps: float):
        if not raw.is_cuda:
            raise RuntimeError("spfsplatv2_b200.adapter needs CUDA tensors (no CPU fallback on the product path)")
        lead = raw.shape[:-1]
        r = raw.detach().float().contiguous().view(-1, raw.shape[-1])
        n = r.shape[0]
        dev = r.device
        scales = torch.empty(n, 3, dtype=torch.float32, device=dev)
        rots = torch.empty(n, 4, dtype=torch.float32, device=dev)
        sh = torch.empty(n, 3, d_sh, dtype=torch.float32, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        L.check(L.lib().spf_adapter_forward(_p(r), n, d_sh, float(eps), _p(scales), _p(rots), _p(sh), stream), "spf_adapter_forward")
        ctx.save_for_backward(r)
        ctx.meta = (d_sh, eps, raw.shape)
        return scales.view(*lead, 3), rots.view(*lead, 4), sh.view(*lead, 3, d_sh)

    @staticmethod
    def backward(ctx, g_scales, g_rots, g_sh):
        (r,) = ctx.saved_tensors
        d_sh, eps, shape = ctx.meta
        n = r.shape[0]
        c = lambda g: None if g is None else g.float().contiguous()
        gs, gr, gh = c(g_scales), c(g_rots), c(g_sh)
        d_raw = torch.empty_like(r)
        stream = C.c_void_p(torch.cuda.current_stream(r.device).cuda_stream)
        L.check(L.lib().spf_adapter_backward(_p(r), _p(gs), _p(gr), _p(gh), n, d_sh, float(eps), _p(d_raw), stream),
                "spf_adapter_backward")
        return d_raw.view(shape), None, None


class _Head(torch.autograd.Function):
    """83-channel head output -> (opacities, scales, rotations, harmonics) in one kernel each way."""

    @staticmethod
    def forward(ctx, raw: Tensor, d_sh: int, eps: float, exponent: float):
        if not raw.is_cuda:
            raise RuntimeError("spfsplatv2_b200.adapter needs CUDA tensors (no CPU fallback on the product path)")
        lead = raw.shape[:-1]
        r = raw.detach().float().contiguous().view(-1, raw.shape[-1])
        n, dev = r.shape[0], r.device
        opac = torch.empty(n, dtype=torch.float32, device=dev)
        scales = torch.empty(n, 3, dtype=torch.float32, device=dev)
        rots = torch.empty(n, 4, dtype=torch.float32, device=dev)
        sh = torch.empty(n, 3, d_sh, dtype=torch.float32, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            L.check(L.lib().spf_head_forward(_p(r), n, d_sh, float(eps), float(exponent), _p(opac), _p(scales), _p(rots), _p(sh),
                                             stream), "spf_head_forward")
        ctx.save_for_backward(r)
        ctx.meta = (d_sh, eps, exponent, raw.shape)
        return opac.view(*lead), scales.view(*lead, 3), rots.view(*lead, 4), sh.view(*lead, 3, d_sh)

    @staticmethod
    def backward(ctx, g_opac, g_scales, g_rots, g_sh):
        (r,) = ctx.saved_tensors
        d_sh, eps, exponent, shape = ctx.meta
        c = lambda g: None if g is None else g.float().contiguous()
        go, gs, gr, gh = c(g_opac), c(g_scales), c(g_rots), c(g_sh)
        d_raw = torch.empty_like(r)
        stream = C.c_void_p(torch.cuda.current_stream(r.device).cuda_stream)
        with torch.cuda.device(r.device):
            L.check(L.lib().spf_head_backward(_p(r), _p(go), _p(gs), _p(gr), _p(gh), r.shape[0], d_sh, float(eps), float(exponent),
                                              _p(d_raw), stream), "spf_head_backward")
        return d_raw.view(shape), None, None, None


@dataclass
class OpacityMappingCfg:
    """encoder.opacity_mapping of the reference (config/model/encoder/spfsplatv2.yaml:6-9)."""
    initial: float = 0.0
    final: float = 0.0
    warm_up: int = 1


def opacity_exponent(cfg: OpacityMappingCfg, global_step: int) -> float:
    """The exponent schedule of EncoderSPFSplatV2.map_pdf_to_opacity (encoder_spfsplatv2.py:153-155)."""
    x = cfg.initial + min(global_step / cfg.warm_up, 1) * (cfg.final - cfg.initial)
    return 2 ** x


@dataclass
class GaussianAdapterCfg:
    gaussian_scale_min: float
    gaussian_scale_max: float
    sh_degree: int


class UnifiedGaussianAdapter(torch.nn.Module):
    """Same call contract as the reference's UnifiedGaussianAdapter.forward(means, opacities, raw_gaussians, eps)."""

    def __init__(self, cfg: GaussianAdapterCfg):
        super().__init__()
        self.cfg = cfg

    @property
    def d_sh(self) -> int:
        return (self.cfg.sh_degree + 1) ** 2

    @property
    def d_in(self) -> int:
        return 7 + 3 * self.d_sh

    def forward(self, means: Tensor, opacities: Tensor, raw_gaussians: Tensor, eps: float = 1e-8) -> Gaussians:
        if raw_gaussians.shape[-1] != self.d_in:
            raise ValueError(f"raw_gaussians has {raw_gaussians.shape[-1]} channels, expected {self.d_in}")
        scales, rotations, sh = _Adapter.apply(raw_gaussians, self.d_sh, eps)
        lead = opacities.shape
        cov = torch.zeros((), dtype=means.dtype, device=means.device).expand(*lead, 3, 3)
        return Gaussians(means=means, covariances=cov, rotations=rotations.expand(*lead, 4), scales=scales.expand(*lead, 3),
                         harmonics=sh.expand(*lead, 3, self.d_sh), opacities=opacities)

    def forward_head(self, means: Tensor, head_out: Tensor, opacity_mapping: OpacityMappingCfg = OpacityMappingCfg(),
                     global_step: int = 0, eps: float = 1e-8) -> Gaussians:
        """The encoder's whole Gaussian post-processing (encoder_spfsplatv2.py:255-268) fused: ``head_out`` [..., 1 + d_in]
        is the Gaussian head's output with the density logit in channel 0; replaces
        ``densities = head_out[..., 0].sigmoid(); opacities = map_pdf_to_opacity(densities, global_step)`` followed by
        ``gaussian_adapter.forward(means, opacities, head_out[..., 1:])`` -- one kernel instead of ~15, no intermediate
        opacity / density tensors, the raw rows read once."""
        if head_out.shape[-1] != self.d_in + 1:
            raise ValueError(f"head output has {head_out.shape[-1]} channels, expected {self.d_in + 1}")
        opac, scales, rotations, sh = _Head.apply(head_out, self.d_sh, eps, opacity_exponent(opacity_mapping, global_step))
        cov = torch.zeros((), dtype=means.dtype, device=means.device).expand(*opac.shape, 3, 3)
        return Gaussians(means=means, covariances=cov, rotations=rotations, scales=scales, harmonics=sh, opacities=opac)
