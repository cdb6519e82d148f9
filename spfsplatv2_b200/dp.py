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

Backend-agnostic (`nccl` on the GPUs, `gloo` in the CPU tests); no renderer code in here.  On NVSwitch machines a
bucket allocated with `GradAllReduce.alloc()` lives in symmetric memory and is reduced INSIDE the switch by this
repo's own multimem kernel (csrc/allreduce.cu, C entry `spf_multimem_allreduce_f32`) instead of NCCL; torch's
symmetric-memory module only provides the allocation, the multicast binding and the cross-rank barrier.
"""
from __future__ import annotations

import ctypes
import os
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


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process (and therefore its pinned-host-memory pages, first-touch) to the CPUs of the NUMA node the
    GPU's PCIe root port hangs off, so that every rank's host->device copies stay on its own socket instead of
    crossing the inter-socket link.  torchrun does not do this.  Call before allocating pinned buffers.
    Returns the node, or None when the topology cannot be read (no GPU, no sysfs, single node) -- never raises."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set(_parse_cpulist(f.read()))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:       # noqa: BLE001
        return None


class GradAllReduce:
    """Sum-all-reduce of a set of gradient tensors, flattened into one bucket, launched on a side stream as soon
    as the producer (projection backward) has been enqueued; `wait()` joins it back into the current stream.

    On CUDA the bucket copy + collective run on `self.stream` after an event recorded on the producer's stream,
    so they overlap whatever the caller enqueues next (the next step's forward).  On CPU (gloo) it is synchronous."""

    def __init__(self, device: torch.device, group: Optional[dist.ProcessGroup] = None, backend: str = "auto",
                 nvls_blocks: Optional[int] = None):
        """backend: "nccl" = torch.distributed only; "nvls" = in-switch multimem kernel for buckets from `alloc()`
        (raises if the machine has no multicast support); "auto" = nvls when available, else nccl.  Overridable
        with SPF_ALLREDUCE=nccl|nvls|auto; SPF_NVLS_BLOCKS sets the CTAs the multimem kernel may occupy."""
        self.device = torch.device(device)
        self.group = group
        self.stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self._bucket: Optional[Tensor] = None
        self._views: List[Tensor] = []
        self._targets: List[Tensor] = []
        self.backend = os.environ.get("SPF_ALLREDUCE", backend)
        if self.backend not in ("auto", "nccl", "nvls"):
            raise ValueError(f"unknown all-reduce backend {self.backend!r}")
        self.nvls_blocks = int(os.environ.get("SPF_NVLS_BLOCKS", nvls_blocks or 16))
        self._symm: Dict[int, tuple] = {}      # data_ptr -> (tensor, symmetric-memory handle)
        self.nvls_error: Optional[str] = None  # why "auto" fell back to nccl, if it did
        # SPF_NVLS_FUSED_BARRIER=1: cross-rank barriers INSIDE the multimem kernel (signal pads) instead of two barrier
        # launches around it.  Correct (sums verified at N=2), but measured slower next to the renderer: while a rank
        # waits for a late peer, all 16 CTAs x 1024 threads of the reduction sit resident on their SMs, where the
        # stand-alone barrier kernel parks one small CTA (2 GPUs, no step dependency: 37.9 k vs 40.0 k views/s,
        # profiles/r2_nvls_barrier_ab.md).  Default: separate barrier launches.  Switched off for every bucket, on every
        # rank together, if the collective self-test fails with it.
        self.fused_barrier = os.environ.get("SPF_NVLS_FUSED_BARRIER", "0") == "1"

    # ---- symmetric-memory buckets (NVLS path) ----------------------------------------------------------------
    def alloc(self, numel: int, dtype: torch.dtype = torch.float32) -> Tensor:
        """A zero-filled gradient bucket.  With the nvls backend (CUDA, world > 1, multicast available) it is
        symmetric memory bound to the group's multicast object and `launch([bucket])` reduces it in the switch;
        otherwise a plain tensor reduced by torch.distributed.  COLLECTIVE: every rank must call it, same sizes."""
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        want = self.backend in ("auto", "nvls") and self.device.type == "cuda" and world > 1 and dtype == torch.float32
        if want:
            t, why = None, None
            try:
                import torch.distributed._symmetric_memory as symm_mem
                padded = (numel + 3) // 4 * 4
                with torch.cuda.device(self.device):
                    t = symm_mem.empty(padded, dtype=dtype, device=self.device)
                    hdl = symm_mem.rendezvous(t, group=self.group if self.group is not None else dist.group.WORLD)
                if not int(hdl.multicast_ptr):
                    raise RuntimeError("symmetric memory has no multicast binding (no NVSwitch multicast support)")
            except Exception as exc:        # noqa: BLE001 -- anything here means "no NVLS on this machine"
                why = f"{type(exc).__name__}: {exc}"
            # The backend must be the same on every rank, and the self-test below is itself collective: agree first
            # that every rank has its bucket, then that every rank saw the right sums.
            if self._all_ranks(why is None):
                self._symm[t.data_ptr()] = (t, hdl)
                for attempt in (0, 1):
                    why = None
                    try:
                        if not self._self_test(t, world):
                            why = "multimem self-test gave wrong sums"
                    except Exception as exc:    # noqa: BLE001
                        why = f"{type(exc).__name__}: {exc}"
                    if self._all_ranks(why is None):
                        t.zero_()
                        return t[:numel]
                    if not self.fused_barrier:
                        break
                    self.fused_barrier = False      # every rank takes this branch together: retry with barrier launches
                    self.nvls_error = f"in-kernel barrier disabled ({why})"
            self._symm.pop(t.data_ptr() if t is not None else 0, None)   # buckets handed out earlier stay on their backend
            why = why or "another rank failed"
            if self.backend == "nvls":
                raise RuntimeError(f"nvls all-reduce unavailable: {why}")
            self.nvls_error = why
        elif self.backend == "nvls" and world > 1:
            raise RuntimeError("nvls all-reduce needs CUDA fp32 buckets")
        return torch.zeros(numel, dtype=dtype, device=self.device)

    def _all_ranks(self, ok: bool) -> bool:
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        return int(flag.item()) == 1

    def _self_test(self, t: Tensor, world: int) -> bool:
        """rank r fills its bucket with r+1 (+ a position ramp); after the reduction every element must hold the sum."""
        rank = dist.get_rank(self.group)
        ramp = (torch.arange(t.numel(), device=self.device, dtype=torch.float32) % 64.0)
        t.copy_(ramp * (rank + 1) + (rank + 1))
        self._launch_nvls(t)
        self.wait()
        tri = world * (world + 1) // 2
        good = bool(torch.equal(t, ramp * tri + tri))
        torch.cuda.synchronize(self.device)
        return good

    def uses_nvls(self, t: Tensor) -> bool:
        return t.data_ptr() in self._symm

    @staticmethod
    def _mc_offset(hdl) -> int:
        # torch >= 2.8 pools symmetric allocations: the handle's multicast address is the pool block's, the tensor
        # sits `offset` bytes into it
        return int(getattr(hdl, "offset", 0) or 0)

    def _launch_nvls(self, g: Tensor) -> None:
        from . import _lib
        t, hdl = self._symm[g.data_ptr()]
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        mc = ctypes.c_void_p(int(hdl.multicast_ptr) + int(self._mc_offset(hdl)))
        pad_words = int(getattr(hdl, "signal_pad_size", 0)) // 4
        # flags live in the upper half of the signal pad: torch's own barrier / signal ops use the low channels
        base = pad_words // 2
        fused = self.fused_barrier and pad_words - base >= 2 * self.nvls_blocks * int(hdl.world_size)
        with torch.cuda.stream(self.stream), torch.cuda.device(self.device):
            if fused:
                _lib.check(_lib.lib().spf_multimem_allreduce_f32_fused(
                    mc, t.numel(), int(hdl.rank), int(hdl.world_size), self.nvls_blocks,
                    ctypes.c_void_p(int(hdl.signal_pad_ptrs_dev)), base, pad_words,
                    ctypes.c_void_p(self.stream.cuda_stream)), "spf_multimem_allreduce_f32_fused")
            else:
                hdl.barrier(channel=0)           # every rank's bucket is final
                _lib.check(_lib.lib().spf_multimem_allreduce_f32(
                    mc, t.numel(), int(hdl.rank), int(hdl.world_size), self.nvls_blocks,
                    ctypes.c_void_p(self.stream.cuda_stream)), "spf_multimem_allreduce_f32")
                hdl.barrier(channel=0)           # every slice has been broadcast to every rank

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
            if self._symm and g.data_ptr() in self._symm:
                self._launch_nvls(g)
                return
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


class NvlsCommHook:
    """DDP communication hook: ``ddp.register_comm_hook(None, NvlsCommHook(reducer))`` replaces the NCCL all-reduce that
    torch DDP issues per gradient bucket (the reference trains with Lightning's "ddp_find_unused_parameters_true"
    strategy, /root/reference/src/main.py:141-145) by this repo's reducer: the bucket is copied into a symmetric-memory
    twin (allocated collectively the first time a bucket of that index and size is seen -- DDP presents buckets in the
    same order on every rank), summed inside the NVSwitch by the multimem kernel on the reducer's side stream, averaged
    over the world size (DDP's contract for the default hook) and copied back.  Falls back to torch.distributed when
    the reducer has no NVLS (CPU / gloo, single node without multicast): same plumbing, same result."""

    def __init__(self, reducer: "GradAllReduce"):
        self.reducer = reducer
        self._twins: Dict[tuple, Tensor] = {}

    def __call__(self, state, bucket) -> "torch.futures.Future[Tensor]":
        buf = bucket.buffer()
        red = self.reducer
        world = dist.get_world_size(red.group)
        fut: torch.futures.Future = torch.futures.Future(devices=[red.device]) if red.device.type == "cuda" else torch.futures.Future()
        key = (bucket.index(), buf.numel(), buf.dtype)
        twin = self._twins.get(key)
        if twin is None:
            twin = red.alloc(buf.numel(), buf.dtype)      # collective
            self._twins[key] = twin
        if red.stream is not None:
            red.stream.wait_stream(torch.cuda.current_stream(red.device))
            with torch.cuda.stream(red.stream):
                twin.copy_(buf)
                red.launch([twin])                        # enqueues on red.stream (already current: wait_stream is a no-op)
                torch.mul(twin, 1.0 / world, out=buf)
                fut.set_result(buf)                       # records the side stream: consumers wait on it, not on the host
        else:
            twin.copy_(buf)
            red.launch([twin])
            torch.mul(twin, 1.0 / world, out=buf)
            fut.set_result(buf)
        return fut


def nvls_comm_hook(reducer: "GradAllReduce"):
    """``ddp.register_comm_hook(None, nvls_comm_hook(reducer))``: a plain function (torch DDP inspects the hook's
    ``__name__`` and its annotations as objects) around an NvlsCommHook."""
    impl = NvlsCommHook(reducer)

    def nvls_allreduce_hook(state, bucket):
        return impl(state, bucket)
    nvls_allreduce_hook.__annotations__ = {"bucket": dist.GradBucket, "return": torch.futures.Future[torch.Tensor]}
    nvls_allreduce_hook.impl = impl
    return nvls_allreduce_hook


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
