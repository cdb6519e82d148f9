"""Seeded synthetic scenes shaped like SPFSplatV2's encoder output.

One Gaussian per context pixel (encoder_spfsplatv2.py:240,296-321 in the
reference), scales following the adapter's law ``min(0.3, 0.001*softplus)``
(gaussian_adapter.py:132-133, "init" regime) or a trained-like log-normal pixel
footprint ("trained" regime), SH degree 4 with the adapter's per-degree mask
(gaussian_adapter.py:42-48), OpenCV camera-to-world extrinsics, normalised
intrinsics fx=fy=0.88 (re10k after the 256^2 crop).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
from torch import Tensor


@dataclass
class Scene:
    means: Tensor        # [b,P,3]
    covariances: Tensor  # [b,P,3,3] (never read by the decoder; kept for the Gaussians record)
    rotations: Tensor    # [b,P,4]
    scales: Tensor       # [b,P,3]
    harmonics: Tensor    # [b,P,3,K]
    opacities: Tensor    # [b,P]
    extrinsics: Tensor   # [b,v,4,4] target cameras (c2w, OpenCV)
    intrinsics: Tensor   # [b,v,3,3] normalised
    near: Tensor         # [b,v]
    far: Tensor          # [b,v]
    image_shape: tuple

    def to(self, device):
        kw = {}
        for k, val in self.__dict__.items():
            kw[k] = val.to(device) if isinstance(val, Tensor) else val
        return Scene(**kw)


def _yaw(deg: float) -> Tensor:
    a = math.radians(deg)
    R = torch.eye(4)
    R[0, 0] = math.cos(a); R[0, 2] = math.sin(a)
    R[2, 0] = -math.sin(a); R[2, 2] = math.cos(a)
    return R


def make_scene(seed: int = 0, v_cxt: int = 2, h: int = 256, w: int = 256, d_sh: int = 25,
               regime: str = "init", n_target: int = 1, grid: tuple | None = None,
               with_cov: bool = False) -> Scene:
    """One scene (b=1).  ``grid=(gh,gw)`` overrides the per-view Gaussian grid
    (C1 uses a 32x32 grid rendered at 64x64)."""
    g = torch.Generator().manual_seed(seed)
    gh, gw = grid if grid is not None else (h, w)
    fx = fy = 0.88
    means, scales = [], []
    for k in range(v_cxt):
        tx = 0.0 if v_cxt == 1 else k / (v_cxt - 1)
        u = (torch.arange(gw, dtype=torch.float32) + 0.5) / gw
        v = (torch.arange(gh, dtype=torch.float32) + 0.5) / gh
        vv, uu = torch.meshgrid(v, u, indexing="ij")
        z = torch.exp(torch.empty(gh, gw).uniform_(math.log(2.0), math.log(20.0), generator=g))
        x = (uu - 0.5) / fx * z + tx
        y = (vv - 0.5) / fy * z
        m = torch.stack([x, y, z], dim=-1).reshape(-1, 3)
        means.append(m)
        n = m.shape[0]
        if regime == "init":
            s = torch.clamp_max(0.001 * torch.nn.functional.softplus(torch.randn(n, 3, generator=g)), 0.3)
        else:
            sig_px = torch.exp(math.log(1.5) + 0.8 * torch.randn(n, 3, generator=g)).clamp(0.3, 12.0)
            s = sig_px * z.reshape(-1, 1) / (fx * w)
        scales.append(s)
    means = torch.cat(means)
    scales = torch.cat(scales)
    P = means.shape[0]
    rot = torch.randn(P, 4, generator=g)
    rot = rot / rot.norm(dim=-1, keepdim=True)
    opac = torch.sigmoid(torch.randn(P, generator=g))
    deg = math.isqrt(d_sh) - 1
    mask = torch.ones(d_sh)
    for d in range(1, deg + 1):
        mask[d * d:(d + 1) * (d + 1)] = 0.1 * 0.25 ** d
    harm = torch.randn(P, 3, d_sh, generator=g) * mask
    ext = []
    for t in range(n_target):
        e = _yaw(3.0 + 2.0 * t)
        e[:3, 3] = torch.tensor([0.5 + 0.1 * t, 0.05, -0.1])
        ext.append(e)
    ext = torch.stack(ext)
    K = torch.tensor([[fx, 0, 0.5], [0, fy, 0.5], [0, 0, 1.0]])
    if with_cov:
        cov = torch.zeros(1, P, 3, 3)
    else:
        cov = torch.zeros(1, 1, 3, 3).expand(1, P, 3, 3)
    return Scene(means[None], cov, rot[None], scales[None], harm[None], opac[None],
                 ext[None], K[None, None].repeat(1, n_target, 1, 1),
                 torch.full((1, n_target), 0.5), torch.full((1, n_target), 500.0), (h, w))


def make_batch(b: int, **kw) -> Scene:
    """b different scenes (seeds seed..seed+b-1) stacked on the batch axis."""
    seed = kw.pop("seed", 0)
    scenes = [make_scene(seed=seed + i, **kw) for i in range(b)]
    def cat(name):
        return torch.cat([getattr(s, name) for s in scenes], dim=0)
    first = scenes[0]
    return Scene(cat("means"), cat("covariances") if first.covariances.is_contiguous() else
                 first.covariances.expand(b, -1, -1, -1), cat("rotations"), cat("scales"),
                 cat("harmonics"), cat("opacities"), cat("extrinsics"), cat("intrinsics"),
                 cat("near"), cat("far"), first.image_shape)
