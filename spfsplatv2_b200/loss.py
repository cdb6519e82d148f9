"""Fused image losses on the decoder output, over ``spf_image_mse``.

Mirrors the reference's interfaces for what consumes ``output.color``:
  * ``LossMse.forward(prediction, image, gaussians, global_step)``  /root/reference/src/loss/loss_mse.py:36-51
  * ``compute_psnr(ground_truth, predicted)``                        /root/reference/src/evaluation/metrics.py:12-19
One pass over the images computes the loss AND dL/dcolor, so the image gradient never round-trips through a chain of
torch elementwise kernels.  CUDA tensors only (no CPU fallback on the product path).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch
from torch import Tensor

from . import _lib as L


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _launch(pred: Tensor, target: Tensor, n_images: int, clip: bool, grad_scale: float, want_grad: bool):
    if not pred.is_cuda:
        raise RuntimeError("spfsplatv2_b200.loss needs CUDA tensors (no CPU fallback on the product path)")
    lib = L.lib()
    pred = pred.detach().float().contiguous()
    target = target.detach().float().contiguous()
    if pred.shape != target.shape:
        target = target.expand_as(pred).contiguous()
    n_per = pred.numel() // n_images
    nb = lib.spf_image_mse_blocks(n_per)
    dev = pred.device
    scratch = torch.empty(n_images * nb + n_images + 1, dtype=torch.float32, device=dev)
    per_img = scratch[n_images * nb:n_images * nb + n_images]
    mean_all = scratch[n_images * nb + n_images:]
    grad = torch.empty_like(pred) if want_grad else None
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    L.check(lib.spf_image_mse(_p(pred), _p(target), n_images, n_per, int(clip), float(grad_scale), _p(grad), _p(scratch),
                              _p(per_img), _p(mean_all), stream), "spf_image_mse")
    return per_img, mean_all, grad


class _MseLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prediction: Tensor, image: Tensor, weight: float):
        n_images = prediction.shape[0] if prediction.dim() > 1 else 1
        want = ctx.needs_input_grad[0]
        per_img, mean_all, grad = _launch(prediction, image, n_images, False, 2.0 * weight / prediction.numel(), want)
        ctx.grad = grad
        ctx.pshape = prediction.shape
        out = mean_all.reshape(())
        return out * weight if weight != 1.0 else out.clone()

    @staticmethod
    def backward(ctx, g):
        if ctx.grad is None:
            return None, None, None
        return (ctx.grad * g).view(ctx.pshape), None, None


def mse_loss(prediction: Tensor, image: Tensor, weight: float = 1.0) -> Tensor:
    """weight * ((prediction - image) ** 2).mean(); differentiable wrt ``prediction`` (the rendered colour)."""
    return _MseLoss.apply(prediction, image, float(weight))


@dataclass
class LossMseCfg:
    weight: float
    apply_after_step: int


class LossMse(torch.nn.Module):
    """Same call contract as the reference's LossMse (loss_mse.py:35-51)."""
    name = "mse"

    def __init__(self, cfg: LossMseCfg):
        super().__init__()
        self.cfg = cfg

    def forward(self, prediction: Tensor, image: Tensor, gaussians=None, global_step: int = 0) -> Tensor:
        if global_step < self.cfg.apply_after_step:
            return torch.tensor(0, dtype=torch.float32, device=image.device)
        return mse_loss(prediction, image, self.cfg.weight)


@torch.no_grad()
def compute_psnr(ground_truth: Tensor, predicted: Tensor) -> Tensor:
    """[batch, c, h, w] x2 -> [batch]; clip both to [0,1], per-image MSE, -10 log10 (metrics.py:12-19)."""
    per_img, _, _ = _launch(predicted, ground_truth, predicted.shape[0], True, 0.0, False)
    return -10 * per_img.clone().log10()
