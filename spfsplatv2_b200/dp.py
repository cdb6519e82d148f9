"""Data-parallel plumbing for the splatting decoder: one process per GPU, views sharded over ranks, ONE gradient
all-reduce per step on a side stream.

What the reference does: Lightning DDP (`strategy="ddp_find_unused_parameters_true"`,
/root/reference/src/main.py:141-145) -- every rank renders its own scenes and torch DDP sum-reduces the
replicated parameters' gradients over NCCL.  The renderer itself has no cross-view state
(/root/reference/src/model/decoder/cuda_splatting.py:96-143 is a plain loop), so the path shards with no
data-path collective; the only exchange is the gradient of whatever is REPLICATED:

  * DDP-faithful (training): different scenes per rank; replicated = upstream (encoder) parameters.  In the
    renderer benchmark a fixed-size stand-in buffer plays that role (bench.py).
  * shared-scene (test-time pose alignment / video, model_wrapper.py:539-590,941-956): the Gaussians are
    replicated, the views are sharded -> the Gaussian-parameter gradients are summed over ranks, the per-view
    pose gradients stay rank-local.

Backend-agnostic (`nccl` on the GPUs, `gloo` in the CPU tests); no renderer code in here.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin partition: rank r owns items {i : i mod world == r}.  Ragged when world does
    not divide n_items; empty when rank >= n_items."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_items, world))


class GradAllReduce:
    """Sum-all-reduce of a set of gradient tensors, flattened into one bucket, launched on a side stream as soon
    as the producer (projection backward) has been enqueued; `wait()` joins it back into the current stream.

    On CUDA the bucket copy + collective run on `self.stream` after an event recorded on the producer's stream,
    so they overlap whatever the caller enqueues next (the next step's forward).  On CPU (gloo) it is synchronous."""

    def __init__(self, device: torch.device, group: Optional[dist.ProcessGroup] = None):
        self.device = torch.device(device)
        self.group = group
        self.stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self._bucket: Optional[Tensor] = None
        self._views: List[Tensor] = []
        self._targets: List[Tensor] = []

    def _ensure_bucket(self, numel: int, dtype: torch.dtype):
        if self._bucket is None or self._bucket.numel() < numel or self._bucket.dtype != dtype:
            self._bucket = torch.empty(numel, dtype=dtype, device=self.device)

    def launch(self, grads: Sequence[Tensor]) -> None:
        grads = [g for g in grads if g is not None]
        if not grads:
            return
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        self._targets = list(grads)
        if world == 1:
            self._views = []
            return
        if len(grads) == 1 and grads[0].is_contiguous():
            # a single tensor IS the bucket: reduce it in place
            g = grads[0]
            self._views = []
            if self.stream is not None:
                self.stream.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(self.stream):
                    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
            else:
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
            return
        total = sum(g.numel() for g in grads)
        self._ensure_bucket(total, grads[0].dtype)
        views, off = [], 0
        for g in grads:
            views.append(self._bucket[off:off + g.numel()].view(g.shape))
            off += g.numel()
        self._views = views
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                torch._foreach_copy_(views, list(grads))
                dist.all_reduce(self._bucket[:total], op=dist.ReduceOp.SUM, group=self.group)
                torch._foreach_copy_(list(grads), views)
        else:
            for v, g in zip(views, grads):
                v.copy_(g)
            dist.all_reduce(self._bucket[:total], op=dist.ReduceOp.SUM, group=self.group)
            for v, g in zip(views, grads):
                g.copy_(v)

    def wait(self) -> None:
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)


def render_views_sharded(render_view: Callable[[int], Tensor], loss_of_view: Callable[[int, Tensor], Tensor],
                         n_views: int, replicated: Dict[str, Tensor], reducer: GradAllReduce) -> Dict[str, Tensor]:
    """Shared-scene variant: every rank holds the same (replicated) differentiable tensors, renders only its own
    views, back-propagates its partial loss, and the replicated tensors' gradients are summed over ranks.
    Returns {name: summed gradient}; the loss value returned under "_loss" is the global sum."""
    rank = dist.get_rank(reducer.group) if dist.is_initialized() else 0
    world = dist.get_world_size(reducer.group) if dist.is_initialized() else 1
    mine = shard_indices(n_views, rank, world)
    first = next(iter(replicated.values()))
    loss = torch.zeros((), dtype=first.dtype, device=first.device)
    for i in mine:
        loss = loss + loss_of_view(i, render_view(i))
    for t in replicated.values():
        t.grad = None
    if mine:
        loss.backward()
    grads = {}
    for k, t in replicated.items():
        grads[k] = t.grad if t.grad is not None else torch.zeros_like(t)
    loss_buf = loss.detach().reshape(1).clone()
    reducer.launch(list(grads.values()) + [loss_buf])
    reducer.wait()
    grads["_loss"] = loss_buf
    return grads
