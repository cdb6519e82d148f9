"""Test-time pose alignment as ONE captured CUDA graph per optimisation step.

The reference's ``ModelWrapper.test_step_align`` (/root/reference/src/model/model_wrapper.py:539-590) refines the target
cameras against the rendered Gaussians: ``pose_align_steps`` iterations of  decoder forward -> image loss -> backward ->
Adam step on the extrinsics, driven from Python (about 25 kernel launches and several host round trips per iteration for
a few hundred microseconds of GPU work).  Here the whole iteration -- fused camera setup, projection, binning, sort,
blend, fused MSE, the backward chain down to dL/dextrinsics and a capturable Adam update -- is captured once and
replayed; the host only enqueues graph launches.

Same contract as the reference loop: the extrinsics start from ``initial_extrinsics``, Adam with ``lr``, the returned
output is the render of the LAST iteration (taken before that iteration's update, as in the reference) and the returned
extrinsics are the updated parameter.  Only capturable losses can be part of the graph: the fused MSE is built in, others
come through ``extra_loss`` (LPIPS stays the caller's business, SURVEY.md §2 out of scope).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
from torch import Tensor

from .loss import mse_loss

import contextlib

_null = contextlib.nullcontext

EAGER_WARMUP = 3          # eager iterations before the capture (they size the rasterizer's data-dependent buffers)


def pose_align(decoder, gaussians, initial_extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
               image_shape: tuple, target_image: Tensor, steps: int, lr: float, mse_weight: float = 1.0,
               extra_loss: Optional[Callable[[Tensor, Tensor], Tensor]] = None, use_graph: bool = True,
               check_every: int = 50):
    """Returns (decoder output of the last iteration, refined extrinsics [b,v,4,4], list of loss tensors seen at the check
    points).  ``gaussians`` are held fixed (detached), as in the reference (the encoder is frozen there).

    A replay whose camera has moved far enough to outgrow the duplicate buffers frozen into the graph yields a NaN loss
    (the blend kernel poisons the image); the loop looks at the loss every ``check_every`` iterations, and on NaN rolls
    the pose and the optimiser state back to the last good check point and finishes eagerly."""
    from .decoder import Gaussians
    dev = initial_extrinsics.device
    g = Gaussians(*[t.detach() for t in (gaussians.means, gaussians.covariances, gaussians.rotations, gaussians.scales,
                                         gaussians.harmonics, gaussians.opacities)])
    graphed = use_graph and dev.type == "cuda"
    # The parameter, the eager warm-up iterations and the capture all live on ONE side stream: autograd accumulates a
    # leaf's gradient on the stream the leaf was created on, and a capture must not touch any other stream.
    side = torch.cuda.Stream(dev) if graphed else None
    if graphed:
        side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side) if graphed else _null():
        extrinsics = torch.nn.Parameter(initial_extrinsics.detach().clone())
    opt = torch.optim.Adam([{"params": [extrinsics], "lr": lr}], capturable=graphed)
    out_box = {}

    def iteration():
        out = decoder.forward(g, extrinsics, intrinsics, near, far, image_shape)
        loss = mse_loss(out.color, target_image, mse_weight)
        if extra_loss is not None:
            loss = loss + extra_loss(out.color, target_image)
        loss.backward()
        opt.step()
        out_box["out"] = out
        return loss

    def eager_iteration():
        from .rasterizer import DuplicateCapacityError
        opt.zero_grad(set_to_none=True)
        try:
            return iteration()
        except DuplicateCapacityError:      # the camera moved enough to outgrow the buffers: capacity raised, go again
            opt.zero_grad(set_to_none=True)
            return iteration()

    losses = []
    done = 0
    n_eager = steps if not graphed else min(steps, EAGER_WARMUP)
    with torch.cuda.stream(side) if graphed else _null():
        for _ in range(n_eager):
            loss = eager_iteration()
            done += 1
    if graphed:
        torch.cuda.current_stream(dev).wait_stream(side)
    if done:
        losses.append(loss.detach())
    if done == steps:
        return out_box["out"], extrinsics.detach(), losses

    def snapshot():
        return extrinsics.detach().clone(), {k: (v.clone() if torch.is_tensor(v) else v) for k, v in opt.state[extrinsics].items()}

    def restore(snap):
        with torch.no_grad():
            extrinsics.copy_(snap[0])
            for k, v in snap[1].items():
                if torch.is_tensor(v):
                    opt.state[extrinsics][k].copy_(v)

    torch.cuda.synchronize(dev)
    good = snapshot()
    graph = torch.cuda.CUDAGraph()
    opt.zero_grad(set_to_none=True)
    with torch.cuda.graph(graph, stream=side):
        static_loss = iteration()
    # (the capture itself does not execute the iteration)
    with torch.cuda.stream(side):
        while done < steps:
            n = min(check_every, steps - done)
            for _ in range(n):
                graph.replay()
            if bool(torch.isfinite(static_loss)):           # one host read per `check_every` iterations
                done += n
                losses.append(static_loss.detach().clone())
                if done < steps:
                    good = snapshot()
                continue
            restore(good)                                   # overflow inside the graph: finish eagerly from the last good state
            for _ in range(steps - done):
                loss = eager_iteration()
            losses.append(loss.detach())
            done = steps
    torch.cuda.current_stream(dev).wait_stream(side)
    return out_box["out"], extrinsics.detach(), losses
