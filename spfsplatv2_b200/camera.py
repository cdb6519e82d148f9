"""Camera-side host glue of the splatting decoder (pure torch, device-agnostic, sync-free).

Mirrors what the reference computes before it enters the rasterizer:
``get_fov`` (/root/reference/src/geometry/projection.py:269-283),
``get_projection_matrix`` (/root/reference/src/model/decoder/cuda_splatting.py:15-42)
and the scale-invariance / transposition steps of ``render_cuda``
(cuda_splatting.py:66-90).  Same values, but batched and without ``.item()``.
"""
from __future__ import annotations

import torch
from torch import Tensor


def get_fov(intrinsics: Tensor) -> Tensor:
    """[B,3,3] normalised intrinsics -> [B,2] (fov_x, fov_y) in radians: the angle between the
    un-projected midpoints of opposite image edges."""
    inv = torch.linalg.inv(intrinsics)
    edges = torch.tensor([[0.0, 0.5, 1.0], [1.0, 0.5, 1.0], [0.5, 0.0, 1.0], [0.5, 1.0, 1.0]],
                         dtype=intrinsics.dtype, device=intrinsics.device)
    rays = torch.einsum("bij,ej->ebi", inv, edges)
    rays = rays / rays.norm(dim=-1, keepdim=True)
    fov_x = (rays[0] * rays[1]).sum(-1).acos()
    fov_y = (rays[2] * rays[3]).sum(-1).acos()
    return torch.stack((fov_x, fov_y), dim=-1)


def get_projection_matrix(near: Tensor, far: Tensor, fov_x: Tensor, fov_y: Tensor) -> Tensor:
    """Frustum -> x,y in (-1,1), z in (0,1), +z forward.  [B] each -> [B,4,4]."""
    tx = (0.5 * fov_x).tan()
    ty = (0.5 * fov_y).tan()
    right = tx * near
    top = ty * near
    out = torch.zeros((near.shape[0], 4, 4), dtype=torch.float32, device=near.device)
    out[:, 0, 0] = 2 * near / (right + right)
    out[:, 1, 1] = 2 * near / (top + top)
    out[:, 0, 2] = (right - right) / (right + right)
    out[:, 1, 2] = (top - top) / (top + top)
    out[:, 3, 2] = 1
    out[:, 2, 2] = far / (far - near)
    out[:, 2, 3] = -(far * near) / (far - near)
    return out


def camera_setup(extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
                 scale_invariant: bool):
    """Returns (viewmatrix [B,4,4], projmatrix [B,4,4], tanfov [B,2], scale [B]) exactly as the
    reference hands them to the rasterizer (both matrices transposed = row-vector convention);
    ``scale`` = 1/near if scale_invariant else 1 (to be applied to means and scales).
    viewmatrix stays differentiable wrt extrinsics (pose gradient)."""
    if scale_invariant:
        scale = 1.0 / near
        extrinsics = extrinsics.clone()
        extrinsics[..., :3, 3] = extrinsics[..., :3, 3] * scale[:, None]
        near = near * scale
        far = far * scale
    else:
        scale = torch.ones_like(near)
    fov = get_fov(intrinsics)
    tanfov = (0.5 * fov).tan()
    proj = get_projection_matrix(near, far, fov[:, 0], fov[:, 1]).transpose(1, 2)
    view = torch.linalg.inv(extrinsics).transpose(1, 2)
    return view, proj, tanfov, scale
