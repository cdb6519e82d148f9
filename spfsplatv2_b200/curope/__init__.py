"""Drop-in module for the reference's ``curope`` extension
(/root/reference/src/model/encoder/backbone/croco/curope/curope.cpp:49-69): ``rope_2d(tokens,
positions, base, fwd)`` rotates ``tokens`` [B,N,H,D] IN PLACE.  Also ``cuRoPE2D`` / ``cuRoPE2D_func``
with the interface of curope2d.py:12-40."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib as L

_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2, torch.float64: 3}


def rope_2d(tokens: torch.Tensor, positions: torch.Tensor, base: float, fwd: float, tokens2: torch.Tensor = None) -> None:
    """``tokens2`` (extension): a second tensor of the same shape, strides and dtype sharing ``positions`` -- the k of
    a self-attention block next to its q (croco/blocks.py:102-104) -- rotated in the same launch."""
    # same argument checks (and messages) as curope.cpp:54-59 / kernels.cu:91-94
    if tokens.dim() != 4:
        raise RuntimeError("tokens must have 4 dimensions")
    if positions.dim() != 3:
        raise RuntimeError("positions must have 3 dimensions")
    if tokens.size(0) != positions.size(0):
        raise RuntimeError("batch size differs between tokens & positions")
    if tokens.size(1) != positions.size(1):
        raise RuntimeError("seq_length differs between tokens & positions")
    if positions.size(2) != 2:
        raise RuntimeError("positions.shape[2] must be equal to 2")
    if tokens.is_cuda != positions.is_cuda:
        raise RuntimeError("tokens and positions are not on the same device")
    if not tokens.is_cuda:
        raise RuntimeError("spfsplatv2_b200.curope: CUDA tensors required (no CPU fallback on the product path)")
    B, N, H, D = tokens.shape
    if not (tokens.stride(3) == 1 and tokens.stride(2) == D):
        raise RuntimeError("tokens are not contiguous")
    if not positions.is_contiguous():
        raise RuntimeError("positions are not contiguous")
    if D % 4 != 0:
        raise RuntimeError("token dim must be multiple of 4")
    if tokens.dtype not in _DTYPES:
        raise RuntimeError(f"unsupported dtype {tokens.dtype}")
    if positions.dtype != torch.int64:
        raise RuntimeError("positions must be int64")
    if tokens2 is not None and (tokens2.shape != tokens.shape or tokens2.stride() != tokens.stride() or
                                tokens2.dtype != tokens.dtype or tokens2.device != tokens.device):
        raise RuntimeError("q and k must have the same shape, strides, dtype and device")
    if tokens.numel() == 0:
        return
    stream = C.c_void_p(torch.cuda.current_stream(tokens.device).cuda_stream)
    with torch.cuda.device(tokens.device):
        if tokens2 is None:
            rc = L.lib().spf_rope2d(C.c_void_p(tokens.data_ptr()), C.c_void_p(positions.data_ptr()), B, N, H, D,
                                    tokens.stride(0), tokens.stride(1), _DTYPES[tokens.dtype], float(base), float(fwd),
                                    stream)
        else:
            rc = L.lib().spf_rope2d_qk(C.c_void_p(tokens.data_ptr()), C.c_void_p(tokens2.data_ptr()),
                                       C.c_void_p(positions.data_ptr()), B, N, H, D, tokens.stride(0), tokens.stride(1),
                                       _DTYPES[tokens.dtype], float(base), float(fwd), stream)
    L.check(rc, "spf_rope2d")


def rope_2d_qk(q: torch.Tensor, k: torch.Tensor, positions: torch.Tensor, base: float, fwd: float) -> None:
    """q and k ([B,N,H,D] views, e.g. of one fused qkv tensor) rotated IN PLACE in one launch."""
    rope_2d(q, positions, base, fwd, tokens2=k)


class cuRoPE2D_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tokens, positions, base, F0=1):
        ctx.save_for_backward(positions)
        ctx.saved_base = base
        ctx.saved_F0 = F0
        rope_2d(tokens, positions, base, F0)
        ctx.mark_dirty(tokens)
        return tokens

    @staticmethod
    def backward(ctx, grad_res):
        (positions,) = ctx.saved_tensors
        if not (grad_res.stride(3) == 1 and grad_res.stride(2) == grad_res.size(3)):
            grad_res = grad_res.contiguous()
        rope_2d(grad_res, positions, ctx.saved_base, -ctx.saved_F0)
        ctx.mark_dirty(grad_res)
        return grad_res, None, None, None


class cuRoPE2D_qkv_func(torch.autograd.Function):
    """q and k slices of the fused qkv tensor [B,N,3,H,D] through one launch, forward and backward (the backward rotates
    both gradient slices by -F0; the v slice passes through).  One in-place input, one output: the autograd engine does
    not allow a Function to modify two views in place."""

    @staticmethod
    def forward(ctx, qkv, positions, base, F0=1):
        if qkv.dim() != 5 or qkv.size(2) != 3:
            raise RuntimeError("qkv must be [B, N, 3, H, D]")
        ctx.save_for_backward(positions)
        ctx.saved_base = base
        ctx.saved_F0 = F0
        rope_2d(qkv[:, :, 0], positions, base, F0, tokens2=qkv[:, :, 1])
        ctx.mark_dirty(qkv)
        return qkv

    @staticmethod
    def backward(ctx, g):
        (positions,) = ctx.saved_tensors
        if not (g.stride(4) == 1 and g.stride(3) == g.size(4)):
            g = g.contiguous()
        rope_2d(g[:, :, 0], positions, ctx.saved_base, -ctx.saved_F0, tokens2=g[:, :, 1])
        ctx.mark_dirty(g)
        return g, None, None, None


class cuRoPE2D(torch.nn.Module):
    def __init__(self, freq=100.0, F0=1.0):
        super().__init__()
        self.base = freq
        self.F0 = F0

    def forward(self, tokens, positions):
        cuRoPE2D_func.apply(tokens.transpose(1, 2), positions, self.base, self.F0)
        return tokens

    def forward_qkv(self, qkv, positions):
        """Fast path for the self-attention call site (croco/blocks.py:97-104).  Instead of
        ``q, k, v = [qkv[:,:,i] ...]; q = rope(q, xpos); k = rope(k, xpos)`` on the transposed views, pass the fused
        projection output as ``qkv`` [B, N, 3, H, D] (before its transpose): its q and k slices are rotated in place by
        ONE launch (cos/sin evaluated once per token and frequency) and the same tensor is returned."""
        return cuRoPE2D_qkv_func.apply(qkv, positions, self.base, self.F0)


class RotaryPositionEmbedding2D(torch.nn.Module):
    """Drop-in for the VGGT backbone's pure-PyTorch 2-D RoPE
    (/root/reference/src/model/encoder/backbone/vggt/layers/rope.py:62-188): same math as ``cuRoPE2D`` but OUT of place,
    ``tokens`` [B, H, N, D] -> new tensor.  Runs on the same sm_100a kernel (no cos/sin tables, no
    ``int(positions.max())`` host sync); ``scaling_factor`` is accepted and, as in the reference, unused."""

    def __init__(self, frequency: float = 100.0, scaling_factor: float = 1.0):
        super().__init__()
        self.base_frequency = frequency
        self.scaling_factor = scaling_factor

    def forward(self, tokens: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
        assert tokens.size(-1) % 2 == 0, "Feature dimension must be even"
        assert positions.ndim == 3 and positions.shape[-1] == 2, "Positions must have shape (batch_size, n_tokens, 2)"
        work = tokens.transpose(1, 2).contiguous()               # [B, N, H, D] private copy (the op is out of place)
        if work.data_ptr() == tokens.data_ptr():
            work = work.clone()
        work = cuRoPE2D_func.apply(work, positions.contiguous(), float(self.base_frequency), 1.0)
        return work.transpose(1, 2)
