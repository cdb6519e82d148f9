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


def orthographic_setup(extrinsics: Tensor, width: Tensor, height: Tensor, near: Tensor, far: Tensor,
                       fov_degrees: float = 0.1):
    """Camera glue of the reference's ``render_cuda_orthographic`` (cuda_splatting.py:173-202): a tiny field of view with
    the camera moved back so that the near plane is `width` wide.  Returns (viewmatrix [B,4,4], projmatrix [B,4,4],
    tanfov [B,2], dump dict with the moved extrinsics / fov / near / far).  Batched where the reference only works
    for B = 1 (it writes a [B] tensor into one matrix element)."""
    b = extrinsics.shape[0]
    dev = extrinsics.device
    fov_x = torch.tensor(fov_degrees, device=dev).deg2rad()
    tan_fov_x = (0.5 * fov_x).tan()
    distance_to_near = (0.5 * width) / tan_fov_x
    tan_fov_y = 0.5 * height / distance_to_near
    fov_y = (2 * tan_fov_y).atan()
    near = near + distance_to_near
    far = far + distance_to_near
    move_back = torch.eye(4, dtype=torch.float32, device=dev).repeat(b, 1, 1)
    move_back[:, 2, 3] = -distance_to_near
    extrinsics = extrinsics @ move_back
    proj = get_projection_matrix(near, far, fov_x.expand(b), fov_y).transpose(1, 2)
    view = torch.linalg.inv(extrinsics).transpose(1, 2)
    tanfov = torch.stack([tan_fov_x.expand(b), tan_fov_y.expand(b)], dim=-1)
    return view, proj, tanfov, dict(extrinsics=extrinsics, fov_x=fov_x, fov_y=fov_y, near=near, far=far)


class _CameraSetupCUDA(torch.autograd.Function):
    """camera_setup as ONE kernel (csrc/camera.cu) with the pose-gradient path extrinsics <- viewmatrix."""

    @staticmethod
    def forward(ctx, extrinsics, intrinsics, near, far, scale_invariant: bool):
        import ctypes as C
        from . import _lib as L
        dev = extrinsics.device
        B = extrinsics.shape[0]
        ext = extrinsics.detach().float().contiguous()
        intr = intrinsics.detach().float().contiguous()
        nr = near.detach().float().contiguous()
        fr = far.detach().float().contiguous()
        view = torch.empty(B, 4, 4, dtype=torch.float32, device=dev)
        proj = torch.empty(B, 4, 4, dtype=torch.float32, device=dev)
        tanfov = torch.empty(B, 2, dtype=torch.float32, device=dev)
        scale = torch.empty(B, dtype=torch.float32, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        L.check(L.lib().spf_camera_forward(B, int(scale_invariant), p(ext), p(intr), p(nr), p(fr), p(view), p(proj),
                                           p(tanfov), p(scale), stream), "spf_camera_forward")
        ctx.save_for_backward(nr, view)
        ctx.scale_invariant = scale_invariant
        ctx.mark_non_differentiable(proj, tanfov, scale)
        return view, proj, tanfov, scale

    @staticmethod
    def backward(ctx, g_view, _gp, _gt, _gs):
        import ctypes as C
        from . import _lib as L
        nr, view = ctx.saved_tensors
        dev = view.device
        B = view.shape[0]
        g = g_view.float().contiguous()
        d_ext = torch.empty(B, 4, 4, dtype=torch.float32, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        L.check(L.lib().spf_camera_backward(B, int(ctx.scale_invariant), p(nr), p(view), p(g), p(d_ext), stream),
                "spf_camera_backward")
        return d_ext, None, None, None, None


def camera_setup_cuda(extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor, scale_invariant: bool):
    """Same outputs as ``camera_setup`` from one kernel launch (CUDA tensors only)."""
    if not extrinsics.is_cuda:
        raise RuntimeError("camera_setup_cuda needs CUDA tensors (no CPU fallback on the product path)")
    return _CameraSetupCUDA.apply(extrinsics, intrinsics, near, far, scale_invariant)
