"""Batched differentiable Gaussian rasterizer: torch.autograd.Function over the C-ABI library.

PyTorch is plumbing here (device memory, streams, autograd graph); all arithmetic runs in
libspfsplat.so.  One call renders B = S*v views (view i reads scene i // v) with one launch
sequence, replacing the reference's per-view Python loop around diff_gauss_pose
(/root/reference/src/model/decoder/cuda_splatting.py:96-143).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import threading
import time
import weakref
from collections import OrderedDict
from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor

from . import _lib as L

TILE = 16
PROJ_THREADS = 128

# Duplicate-buffer capacity policy: no device->host wait before the kernels are enqueued.  The buffers
# are sized from a per-shape high-water mark; the exact count N comes back through pinned memory while
# the blend kernel is already running, and the forward is re-run (rare) if N exceeded the capacity.
class DuplicateCapacityError(RuntimeError):
    """A forward whose duplicate count was checked late (training call) had more (Gaussian, tile) duplicates than the
    buffers sized from earlier calls of the shape.  Its image was poisoned with NaN by the blend kernel (never a
    plausible wrong picture); the capacity has been raised: run the step again."""


class _Hints(OrderedDict):
    """Per-shape sizing hints: bounded (least recently used shape evicted) and guarded by the module lock."""
    MAX = 64

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        self.move_to_end(key)
        while len(self) > self.MAX:
            self.popitem(last=False)


_lock = threading.RLock()       # guards the sizing hints, the ticket counter and the pinned pools (host-side state only)
_capacity_hint: dict = _Hints()
GROWTH = 1.5
_host_counters: dict = {}     # device index -> (pinned int32[2] tensor, numpy view)
_ticket = [0]
# Pair-log sizing (records per warp): per-shape high-water mark of control[2], fed back asynchronously after each
# backward (pinned copy + event, polled -- never waited on -- by the next forward of that shape).
_pair_cap_hint: dict = _Hints()
_pair_stat: dict = _Hints()         # key -> [[pinned int32[1], event, age in forwards], ...]
_pin_pool: list = []          # recycled (pinned int32[1], event) pairs: cudaHostAlloc per step would cost ~100 us
PAIR_CAP_DEFAULT = 512        # pairs per warp; the log holds 8 bytes per pair (4 KB per warp, 134 MB for 16 views at 256^2)
PAIR_BYTES = 8
PAIR_LOG_BUDGET = 2 << 30     # bytes; above this the capacity is clipped and the densest tiles fall back to recomputation


def _ptr(t: Optional[Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)



def _f32c(t: Tensor) -> Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


@dataclass
class RasterSettings:
    image_height: int
    image_width: int
    sh_degree: int
    scale_modifier: float = 1.0
    views_per_scene: int = 1
    sh_layout_ck: bool = False          # shs given as [S,P,3,K] (encoder-native) instead of [S,P,K,3]
    enable_cov_grad: bool = True
    enable_sh_grad: bool = True
    quat_xyzw: bool = False
    want_alpha: bool = False
    want_means2d_grad: bool = False
    no_tma: bool = False
    bwd_v1: bool = False                # debugging: first-generation blend backward
    pair_log: bool = True               # forward logs contributing pairs for the backward (when grads are needed)
    sync_count: bool = False            # always wait for the duplicate count in the forward (never defer the check)

    def flags(self) -> int:
        f = 0
        if self.sh_layout_ck:
            f |= L.SPF_FLAG_SH_LAYOUT_CK
        if not self.enable_cov_grad:
            f |= L.SPF_FLAG_NO_COV_GRAD
        if not self.enable_sh_grad:
            f |= L.SPF_FLAG_NO_SH_GRAD
        if self.quat_xyzw:
            f |= L.SPF_FLAG_QUAT_XYZW
        if self.no_tma or os.environ.get("SPF_NO_TMA") == "1":
            f |= L.SPF_FLAG_NO_TMA
        if self.bwd_v1 or os.environ.get("SPF_BWD_V1") == "1":
            f |= L.SPF_FLAG_BWD_V1
        return f


class _Workspace:
    """All forward intermediates in ONE device allocation (one caching-allocator call instead of ~20); raw pointers are
    base + offset, tensor views are only materialised on demand (tests / inspection)."""
    _SIZES = {torch.float32: 4, torch.int32: 4, torch.int64: 8}

    def __init__(self, device, spec):
        self.layout = {}
        off = 0
        for name, shape, dtype in spec:
            n = 1
            for d in shape:
                n *= int(d)
            self.layout[name] = (off, tuple(int(d) for d in shape), dtype, n)
            off += (n * self._SIZES[dtype] + 255) // 256 * 256
        self.buf = torch.empty(max(off, 256), dtype=torch.uint8, device=device)
        self.base = self.buf.data_ptr()
        self._views = {}

    def ptr(self, name):
        ent = self.layout.get(name)
        return None if ent is None else C.c_void_p(self.base + ent[0])

    def __contains__(self, name):
        return name in self.layout

    def get(self, name, default=None):
        return self[name] if name in self.layout else default

    def __getitem__(self, name):
        v = self._views.get(name)
        if v is None:
            off, shape, dtype, n = self.layout[name]
            v = self.buf[off:off + n * self._SIZES[dtype]].view(dtype).view(shape)
            self._views[name] = v
        return v


class _Tensors:
    """dict-like view over the state's workspaces (fixed-size one + capacity-dependent one)."""

    def __init__(self, *spaces):
        self.spaces = spaces

    def __contains__(self, name):
        return any(name in w for w in self.spaces)

    def __getitem__(self, name):
        for w in self.spaces:
            if name in w:
                return w[name]
        raise KeyError(name)

    def ptr(self, name):
        for w in self.spaces:
            if name in w:
                return w.ptr(name)
        return None


_STATE_FIELDS = ("xy", "depth", "conic_opacity", "rgb", "radii", "tiles_touched", "dup_offset", "control", "bucket",
                 "slab", "cullbox", "tile_ranges", "final_T", "n_contrib", "accum", "pair_log", "pair_count")

N_COUNTER_SLOTS = 16


def _counters(dev):
    """Ring of {N, ticket} slots in mapped pinned host memory (one slot per in-flight forward)."""
    hc = _host_counters.get(dev.index)
    if hc is None:
        t = torch.zeros(N_COUNTER_SLOTS, 2, dtype=torch.int32).pin_memory()
        hc = (t, t.numpy())
        _host_counters[dev.index] = hc
    return hc


class _State:
    """Forward intermediates kept for backward / inspection (not autograd-tracked)."""
    __slots__ = ("desc", "cin", "cstate", "cout", "keep", "_n_dups", "capacity", "tensors", "key", "ticket", "slot",
                 "dev", "settings", "scratch", "captured", "__weakref__")

    @property
    def n_dups(self) -> int:
        """Duplicate count of this forward.  Read from the pinned slot the scan kernel wrote; normally long there by
        the time anyone asks (the polling loop only waits if the GPU has not reached the scan kernel yet)."""
        if self._n_dups is None:
            self._n_dups = _wait_count(self.dev, self.slot, self.ticket, self.tensors)
        return self._n_dups


def _poll_count(dev, slot, ticket):
    """Non-blocking: the duplicate count of forward `ticket` if the scan kernel has already reported it, else None."""
    _, host_np = _counters(dev)
    if int(host_np[slot, 1]) == ticket:
        return int(host_np[slot, 0])
    return None


def _wait_count(dev, slot, ticket, tensors) -> int:
    """Waits for the scan kernel of forward `ticket` to report N through the pinned slot: a short busy poll (the usual
    wait is a few microseconds), then sleeps of growing length so a long queue ahead of the kernel does not burn a core."""
    _, host_np = _counters(dev)
    spins = 0
    t0 = None
    while True:
        cur = int(host_np[slot, 1])
        if cur == ticket:
            return int(host_np[slot, 0])
        if cur > ticket and (cur - ticket) % N_COUNTER_SLOTS == 0:
            # slot recycled by a later forward: ask the device (rare)
            return None if tensors is None else int(tensors["control"][0])
        spins += 1
        if spins > 2000:
            if t0 is None:
                t0 = time.monotonic()
            waited = time.monotonic() - t0
            time.sleep(min(1e-3, 2e-5 * (1 + (spins - 2000) // 50)))
            if waited > 2.0:
                torch.cuda.current_stream(dev).synchronize()   # surfaces a launch failure instead of hanging
                if int(host_np[slot, 1]) != ticket:
                    raise RuntimeError("spf_raster_forward: duplicate count never arrived (kernel failure?)")


PAIR_FEEDBACK_LAG = 2   # forwards between a backward's report and its use


def _pair_capacity(key, n_warps: int) -> int:
    """Records per warp for the next forward of this shape.  The capacity only ever GROWS, and only when a backward of
    this shape reported an overflow.  The report (a 4-byte pinned copy issued after that backward) is applied at a
    deterministic point -- the PAIR_FEEDBACK_LAG-th later forward of the shape -- so runs are reproducible, and by
    then it has long landed, so the host never stalls the stream it is feeding."""
    cap = int(_pair_cap_hint.get(key, PAIR_CAP_DEFAULT))
    pending = _pair_stat.get(key)
    if pending:
        for ent in pending:
            ent[2] += 1
        while pending and pending[0][2] >= PAIR_FEEDBACK_LAG:
            pin, ev, _ = pending.pop(0)
            ev.synchronize()
            need = int(pin[0])
            _pin_pool.append((pin, ev))
            if need > cap:
                cap = ((int(need * 1.25) + 63) // 64) * 64
                _pair_cap_hint[key] = cap
    return max(64, min(cap, (PAIR_LOG_BUDGET // (PAIR_BYTES * max(n_warps, 1))) // 64 * 64))


_unverified: dict = _Hints()     # key -> state of the last forward whose duplicate count has not been looked at yet


def _check_unverified(key):
    """A training-mode forward whose count was deferred and whose backward never ran (evaluation under enable_grad,
    a discarded pose-search step): look at its count now -- it has long landed -- so that an overflow is reported
    (and the capacity raised) instead of going unnoticed.  Holds (device, slot, ticket, capacity) only, never the
    forward's buffers."""
    prev = _unverified.pop(key, None)
    if prev is None:
        return
    dev, slot, ticket, capacity = prev
    n = _wait_count(dev, slot, ticket, None)
    if n is None:          # the pinned slot was recycled by 16 later forwards: nothing left to look at
        return
    with _lock:
        _capacity_hint[key] = max(int(_capacity_hint.get(key, 0)), int(n * GROWTH) + 1024)
    if n > capacity:
        raise DuplicateCapacityError(
            f"spfsplatv2_b200: the previous forward of this shape had {n} tile duplicates for a buffer capacity of "
            f"{capacity}; its image was NaN-poisoned.  The capacity has been raised: run it again.")


def _launch_forward(st: "_State", fixed: _Workspace, cap: int, color, depth, alpha):
    """(Re)allocate the capacity-dependent buffers for `cap` duplicates and enqueue the whole forward sequence."""
    lib = L.lib()
    dev = st.dev
    capws = _Workspace(dev, [("bucket", (cap,), torch.int64), ("slab", (cap, 12), torch.float32),
                             ("cullbox", (cap, 4), torch.float32)])
    st.tensors = _Tensors(fixed, capws)
    host_t, _ = _counters(dev)
    with _lock:
        _ticket[0] = (_ticket[0] % 0x3fffffff) + 1
        st.ticket = _ticket[0]
    st.slot = st.ticket % N_COUNTER_SLOTS
    st.desc.dup_capacity = cap
    st.desc.ticket = st.ticket
    st.capacity = cap
    st._n_dups = None
    ptrs = [st.tensors.ptr(k) for k in _STATE_FIELDS]
    st.cstate = L.SpfRasterState(*ptrs, None if getattr(st, "captured", False)
                                 else C.c_void_p(host_t.data_ptr() + 8 * st.slot))
    st.cout = L.SpfRasterOut(_ptr(color), _ptr(depth), _ptr(alpha))
    with torch.cuda.device(dev):       # the C entry points launch on the CURRENT device's context
        L.check(lib.spf_raster_forward(C.byref(st.desc), C.byref(st.cin), C.byref(st.cstate), C.byref(st.cout),
                                       _stream(dev)), "spf_raster_forward")


def _forward_impl(s: RasterSettings, means, scales, rots, opac, shs, colors, viewmat, projmat, tanfov, bg,
                  pre_scale, pair_log: bool = False, defer_count: bool = False, raw=None):
    """``raw`` = (head rows [S,P,R], has_density, eps, opacity_exponent) selects the raw-head input of the projection
    kernels (SpfRasterIn.raw_head); scales / rots / shs / colors are then None (and opac too when the rows carry the
    density logit)."""
    lib = L.lib()
    dev = means.device
    if dev.type != "cuda":
        raise RuntimeError("spfsplatv2_b200 rasterizer needs CUDA tensors (no CPU fallback)")
    S, P = means.shape[0], means.shape[1]
    v = s.views_per_scene
    B = S * v
    H, W = s.image_height, s.image_width
    if viewmat.shape[0] != B:
        raise ValueError(f"viewmatrix has {viewmat.shape[0]} views, expected n_scenes*views_per_scene={B}")
    T = ((W + TILE - 1) // TILE) * ((H + TILE - 1) // TILE)
    use_sh = shs is not None
    K = 0
    if raw is not None:
        raw_head, has_density, raw_eps, exponent = raw
        R = raw_head.shape[-1]
        K = (R - (1 if has_density else 0) - 7) // 3
        if raw_head.shape[:2] != (S, P) or R != (1 if has_density else 0) + 7 + 3 * K or K < (s.sh_degree + 1) ** 2:
            raise ValueError(f"head rows {tuple(raw_head.shape)} do not match means {tuple(means.shape)} / sh_degree {s.sh_degree}")
    elif use_sh == (colors is not None):
        raise Exception("Please provide excatly one of either SHs or precomputed colors!")
    elif use_sh:
        K = shs.shape[-1] if s.sh_layout_ck else shs.shape[-2]

    f32, i32 = torch.float32, torch.int32
    key = (dev.index, S, v, P, H, W)
    capturing = torch.cuda.is_current_stream_capturing()
    if not capturing:
        _check_unverified(key)
    hint = _capacity_hint.get(key)
    cap = max(int(hint if hint is not None else 2 * B * P), 1024)
    # CUDA-graph capture (torch.cuda.graph around a whole fwd+loss+bwd step): nothing here may wait on the GPU, and the
    # sizes are frozen into the graph.  Run at least one eager step of the same shape first (it establishes the
    # capacities); afterwards `graph_overflowed(state)` tells whether a replay ever outgrew them.
    if capturing and hint is None:
        raise RuntimeError("spfsplatv2_b200: run one eager forward of this shape before capturing it in a CUDA graph "
                           "(the duplicate-buffer capacity is learned from it)")

    if pair_log and s.pair_log:
        if capturing:
            pair_cap = max(64, min(int(_pair_cap_hint.get(key, PAIR_CAP_DEFAULT)),
                                   (PAIR_LOG_BUDGET // (PAIR_BYTES * max(B * T * 8, 1))) // 64 * 64))
        else:
            pair_cap = _pair_capacity(key, B * T * 8)
    else:
        pair_cap = 0
    st = _State()
    st.dev, st.key, st.settings, st.captured = dev, key, s, capturing
    st.desc = L.SpfRasterDesc(S, v, P, H, W, s.sh_degree, s.flags(), float(s.scale_modifier), cap, 0, pair_cap)
    n_ctrl = lib.spf_raster_control_ints(C.byref(st.desc))
    if n_ctrl < 0:
        L.check(-1, "spf_raster_control_ints")
    st.cin = L.SpfRasterIn(_ptr(means), _ptr(scales), _ptr(rots), _ptr(opac), _ptr(shs), _ptr(colors), K,
                           _ptr(viewmat), _ptr(projmat), _ptr(tanfov), _ptr(bg), _ptr(pre_scale))
    st.keep = (means, scales, rots, opac, shs, colors, viewmat, projmat, tanfov, bg, pre_scale)
    if raw is not None:
        st.cin.raw_head = _ptr(raw_head)
        st.cin.raw_stride, st.cin.raw_has_density = R, int(bool(has_density))
        st.cin.raw_eps, st.cin.opacity_exponent = float(raw_eps), float(exponent)
        st.keep = st.keep + (raw_head,)
    spec = [("xy", (B, P, 2), f32), ("depth", (B, P), f32), ("conic_opacity", (B, P, 4), f32), ("rgb", (B, P, 3), f32),
            ("radii", (B, P), i32), ("tiles_touched", (B, P), i32), ("dup_offset", (B, P), i32), ("control", (n_ctrl,), i32),
            ("tile_ranges", (B * T, 2), i32), ("final_T", (B, H, W), f32), ("n_contrib", (B, H, W), i32),
            ("accum", (B, H, W, 4), f32)]
    if pair_cap > 0:
        spec += [("pair_log", (B * T * 8, pair_cap, PAIR_BYTES // 4), f32), ("pair_count", (B * T * 8,), i32)]
    fixed = _Workspace(dev, spec)
    color = torch.empty(B, 3, H, W, dtype=f32, device=dev)
    depth = torch.empty(B, 1, H, W, dtype=f32, device=dev)
    alpha = torch.empty(B, 1, H, W, dtype=f32, device=dev) if s.want_alpha else None
    _launch_forward(st, fixed, cap, color, depth, alpha)
    _last_state[0] = weakref.ref(st)
    if capturing:
        st._n_dups = cap            # sizes frozen at capture time
        return color, depth, alpha, fixed["radii"], st
    n = None
    if defer_count and hint is not None and not s.sync_count:
        # Steady-state training call: the duplicate count is NOT awaited here (that would make the host wait for the
        # GPU to reach this call's scan kernel on every step).  Capacity = GROWTH x the largest count seen so far for
        # this shape.  If the count has already landed it is checked now (and the forward re-run if needed); otherwise
        # it is verified when the backward starts (see _Rasterize.backward) or by the next forward of the shape.  A
        # forward that did overflow never yields a plausible image: the blend kernel poisons it with NaN.
        n = _poll_count(dev, st.slot, st.ticket)
        if n is None:
            _unverified[key] = (dev, st.slot, st.ticket, cap)
            return color, depth, alpha, fixed["radii"], st
    while True:
        if n is None:
            n = st.n_dups        # polls the pinned slot: the rest of the forward is already queued behind the scan
        st._n_dups = n
        if n <= cap:
            break
        cap = int(n * 1.05) + 1024
        _launch_forward(st, fixed, cap, color, depth, alpha)
        n = None
    with _lock:
        _capacity_hint[key] = max(int(_capacity_hint.get(key, 0)), int(n * GROWTH) + 1024)
    return color, depth, alpha, fixed["radii"], st


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, settings: RasterSettings, means, scales, rots, opac, shs, colors, viewmat, projmat, tanfov,
                bg, pre_scale, means2d):
        args = [None if a is None else _f32c(a.detach()) for a in
                (means, scales, rots, opac, shs, colors, viewmat, projmat, tanfov, bg, pre_scale)]
        need_grad = any(ctx.needs_input_grad)
        color, depth, alpha, radii, st = _forward_impl(settings, *args, pair_log=need_grad, defer_count=need_grad)
        ctx.settings = settings
        ctx.st = st
        ctx.shapes = (means.shape, scales.shape, rots.shape, opac.shape,
                      None if shs is None else shs.shape, None if colors is None else colors.shape,
                      viewmat.shape, None if means2d is None else means2d.shape)
        ctx.dtypes = tuple(None if t is None else t.dtype for t in (means, scales, rots, opac, shs, colors, viewmat, means2d))
        ctx.mark_non_differentiable(radii)
        if alpha is None:
            alpha = torch.empty(0, device=color.device)
            ctx.mark_non_differentiable(alpha)
        return color, depth, alpha, radii

    @staticmethod
    def backward(ctx, g_color, g_depth, g_alpha, _g_radii):
        sh = ctx.shapes
        g = _backward_impl(ctx.st, ctx.settings, g_color, g_depth, g_alpha, want_m2d=sh[7] is not None)
        dt = ctx.dtypes      # gradients are computed in fp32; handed back in each input's own dtype (a no-op for fp32 inputs)
        v = lambda t, i: None if t is None else t.view(sh[i]).to(dt[i])
        return (None, v(g["means"], 0), v(g["scales"], 1), v(g["rots"], 2), v(g["opac"], 3), v(g["shs"], 4), v(g["cols"], 5),
                v(g["view"], 6), None, None, None, None, v(g["m2d"], 7))


def _backward_impl(st: "_State", s: RasterSettings, g_color, g_depth, g_alpha, want_m2d: bool) -> dict:
    """Deferred duplicate-count check, gradient buffers, spf_raster_backward, pair-log sizing feedback.  Returns the
    gradient tensors by name (``raw`` instead of scales / rots / shs for a raw-head forward)."""
    lib = L.lib()
    means, scales, rots, opac, shs, colors = st.keep[:6]
    raw_head = st.keep[11] if len(st.keep) > 11 else None
    dev = means.device
    S, P = means.shape[0], means.shape[1]
    B = S * s.views_per_scene
    NB = (P + PROJ_THREADS - 1) // PROJ_THREADS
    f32 = dict(dtype=torch.float32, device=dev)
    # deferred duplicate-count check (the forward did not wait for it; by now the count has long landed)
    n = st.n_dups
    pend = _unverified.get(st.key)
    if pend is not None and pend[2] == st.ticket:
        _unverified.pop(st.key, None)
    if not st.captured:
        with _lock:
            _capacity_hint[st.key] = max(int(_capacity_hint.get(st.key, 0)), int(n * GROWTH) + 1024)
        if n > st.capacity:
            raise DuplicateCapacityError(
                f"spfsplatv2_b200: {n} tile duplicates exceeded the buffer capacity {st.capacity} chosen from earlier "
                "calls of this shape; the forward image of this step was NaN-poisoned (so was any loss computed from "
                "it).  The capacity has been raised: run the step again.")
    gc = None if g_color is None else _f32c(g_color)
    gd = None if g_depth is None else _f32c(g_depth)
    ga = None if (g_alpha is None or not s.want_alpha) else _f32c(g_alpha)
    gout = L.SpfRasterGradOut(_ptr(gc), _ptr(gd), _ptr(ga))
    scratch = _Workspace(dev, [("dup_grad", (max(n, 1), 12), torch.float32), ("pose_partial", (B, NB, 16), torch.float32)])
    e = lambda t: None if t is None else torch.empty_like(t)
    g = dict(means=e(means), scales=e(scales), rots=e(rots), opac=e(opac), shs=e(shs), cols=e(colors), raw=e(raw_head),
             view=torch.empty(B, 16, **f32), m2d=torch.empty(B, P, 3, **f32) if (s.want_means2d_grad and want_m2d) else None)
    gin = L.SpfRasterGradIn(scratch.ptr("dup_grad"), scratch.ptr("pose_partial"), _ptr(g["means"]), _ptr(g["scales"]),
                            _ptr(g["rots"]), _ptr(g["opac"]), _ptr(g["shs"]), _ptr(g["cols"]), _ptr(g["view"]), _ptr(g["m2d"]),
                            _ptr(g["raw"]))
    with torch.cuda.device(dev):
        L.check(lib.spf_raster_backward(C.byref(st.desc), C.byref(st.cin), C.byref(st.cstate), C.byref(gout),
                                        C.byref(gin), _stream(dev)), "spf_raster_backward")
    if st.desc.pair_capacity > 0 and not st.captured and len(_pair_stat.setdefault(st.key, [])) < 4:
        # feed the largest per-warp pair count back to the sizing of the next forward (no wait)
        pin, ev = _pin_pool.pop() if _pin_pool else (torch.empty(1, dtype=torch.int32).pin_memory(), torch.cuda.Event())
        pin.copy_(st.tensors["control"][2:3], non_blocking=True)
        ev.record(torch.cuda.current_stream(dev))
        _pair_stat[st.key].append([pin, ev, 0])
    return g


class _RasterizeHead(torch.autograd.Function):
    """Raw-head input: the encoder head's rows go straight into the projection kernels."""

    @staticmethod
    def forward(ctx, settings: RasterSettings, means, head, opac, viewmat, projmat, tanfov, bg, pre_scale, means2d,
                eps: float, exponent: float):
        a = [None if t is None else _f32c(t.detach()) for t in (means, opac, viewmat, projmat, tanfov, bg, pre_scale, head)]
        need_grad = any(ctx.needs_input_grad)
        color, depth, alpha, radii, st = _forward_impl(settings, a[0], None, None, a[1], None, None, a[2], a[3], a[4], a[5],
                                                       a[6], pair_log=need_grad, defer_count=need_grad,
                                                       raw=(a[7], opac is None, eps, exponent))
        ctx.settings, ctx.st = settings, st
        ctx.shapes = (means.shape, head.shape, None if opac is None else opac.shape, viewmat.shape,
                      None if means2d is None else means2d.shape)
        ctx.dtypes = tuple(None if t is None else t.dtype for t in (means, head, opac, viewmat, means2d))
        ctx.mark_non_differentiable(radii)
        if alpha is None:
            alpha = torch.empty(0, device=color.device)
            ctx.mark_non_differentiable(alpha)
        return color, depth, alpha, radii

    @staticmethod
    def backward(ctx, g_color, g_depth, g_alpha, _g_radii):
        sh = ctx.shapes
        g = _backward_impl(ctx.st, ctx.settings, g_color, g_depth, g_alpha, want_m2d=sh[4] is not None)
        dt = ctx.dtypes
        v = lambda t, i: None if t is None else t.view(sh[i]).to(dt[i])
        return (None, v(g["means"], 0), v(g["raw"], 1), v(g["opac"], 2), v(g["view"], 3), None, None, None, None, v(g["m2d"], 4),
                None, None)


_last_state = [None]


def last_forward_state():
    """State of the most recent forward issued by this process (None once it has been garbage-collected): lets a caller
    of the autograd entry points reach `graph_overflowed` for a step captured in a CUDA graph."""
    ref = _last_state[0]
    return None if ref is None else ref()


def graph_overflowed(st: "_State") -> bool:
    """After replaying a captured step: did the LAST replay outgrow the duplicate buffer frozen into the graph?  Its
    image (and any loss computed from it) is then NaN -- the blend kernel poisons it -- its gradients are invalid, and
    nothing was written out of bounds: re-capture after an eager step of the denser scene.  Synchronises."""
    ctrl = st.tensors["control"][:4].cpu()
    return bool(int(ctrl[0]) > st.capacity or int(ctrl[1]) != 0)


def rasterize_batched(settings: RasterSettings, means: Tensor, scales: Tensor, rotations: Tensor,
                      opacities: Tensor, shs: Optional[Tensor], colors: Optional[Tensor], viewmatrix: Tensor,
                      projmatrix: Tensor, tanfov: Tensor, bg: Tensor, pre_scale: Optional[Tensor] = None,
                      means2d: Optional[Tensor] = None):
    """means [S,P,3], scales [S,P,3], rotations [S,P,4], opacities [S,P], shs [S,P,K,3] (or [S,P,3,K] with
    settings.sh_layout_ck) or colors [S,P,3]; viewmatrix/projmatrix [B,4,4] (row-vector convention),
    tanfov [B,2], bg [B,3], pre_scale [B] or None.  Returns (color [B,3,H,W], depth [B,1,H,W],
    alpha [B,1,H,W] or empty, radii [B,P] int32).  Differentiable wrt means, scales, rotations,
    opacities, shs/colors and viewmatrix."""
    return _Rasterize.apply(settings, means, scales, rotations, opacities, shs, colors, viewmatrix, projmatrix,
                            tanfov, bg, pre_scale, means2d)


def rasterize_batched_head(settings: RasterSettings, means: Tensor, head_out: Tensor, viewmatrix: Tensor,
                           projmatrix: Tensor, tanfov: Tensor, bg: Tensor, pre_scale: Optional[Tensor] = None,
                           means2d: Optional[Tensor] = None, *, opacities: Optional[Tensor] = None, eps: float = 1e-8,
                           opacity_exponent: float = 1.0):
    """Renders straight from the encoder head's rows: ``head_out`` [S,P,1+7+3K] = [density logit, 3 scale logits,
    4 quaternion components, 3 x K SH coefficients] (or [S,P,7+3K] with ``opacities`` [S,P] given separately, the
    adapter's own contract).  The adapter's maps and the opacity mapping (gaussian_adapter.py:122-150,
    encoder_spfsplatv2.py:146-159,255-268) run inside the projection kernels; scales / rotations / harmonics / opacities
    and their gradients never exist in HBM.  Same outputs as ``rasterize_batched``; differentiable wrt means, head_out,
    opacities and viewmatrix.  Needs P % 4 == 0."""
    return _RasterizeHead.apply(settings, means, head_out, opacities, viewmatrix, projmatrix, tanfov, bg, pre_scale,
                                means2d, float(eps), float(opacity_exponent))


def forward_with_state(settings: RasterSettings, means, scales, rotations, opacities, shs, colors, viewmatrix,
                       projmatrix, tanfov, bg, pre_scale=None, pair_log: bool = False, raw=None):
    """No-autograd forward that also returns the intermediate state (for parity tests / profiling).  ``raw`` =
    (head rows, has_density, eps, opacity_exponent) selects the raw-head input (scales / rotations / shs then None)."""
    args = [None if a is None else _f32c(a.detach()) for a in
            (means, scales, rotations, opacities, shs, colors, viewmatrix, projmatrix, tanfov, bg, pre_scale)]
    if raw is not None:
        raw = (_f32c(raw[0].detach()),) + tuple(raw[1:])
    return _forward_impl(settings, *args, pair_log=pair_log, raw=raw)


def unpack_sorted(st: _State):
    """(point_list int32 [N], keys int64 [N]) of the sorted duplicate list -- parity helper."""
    n = st.n_dups
    dev = st.tensors["slab"].device
    pl = torch.empty(n, dtype=torch.int32, device=dev)
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    L.check(L.lib().spf_raster_unpack_sorted(C.byref(st.desc), C.byref(st.cstate), n, _ptr(pl), _ptr(keys),
                                             _stream(dev)), "spf_raster_unpack_sorted")
    return pl, keys


FWD_STAGES = ("clear_control", "project_forward", "scan", "emit", "tile_sort_pack", "blend_forward")
BWD_STAGES = ("blend_backward", "project_backward", "pose_reduce")


def profile_stages(settings: RasterSettings, means, scales, rotations, opacities, shs, colors, viewmatrix,
                   projmatrix, tanfov, bg, pre_scale, g_color, g_depth, iters: int = 10, raw=None) -> dict:
    """Per-kernel device times (ms, mean over ``iters``) measured with CUDA events around each stage of
    the forward and backward launch sequences (spf_raster_{forward,backward}_stages), on the current
    stream.  Used by bench.py for the roofline of the dominant kernel."""
    lib = L.lib()
    color, depth, alpha, radii, st = forward_with_state(settings, means, scales, rotations, opacities, shs, colors,
                                                        viewmatrix, projmatrix, tanfov, bg, pre_scale, pair_log=True, raw=raw)
    means_c, scales_c, rots_c, opac_c, shs_c, cols_c = st.keep[:6]
    raw_c = st.keep[11] if len(st.keep) > 11 else None
    dev = means_c.device
    S, P = means_c.shape[0], means_c.shape[1]
    B = S * settings.views_per_scene
    NB = (P + PROJ_THREADS - 1) // PROJ_THREADS
    f32 = dict(dtype=torch.float32, device=dev)
    gc, gd = _f32c(g_color), (None if g_depth is None else _f32c(g_depth))
    gout = L.SpfRasterGradOut(_ptr(gc), _ptr(gd), None)
    e = lambda t: None if t is None else torch.empty_like(t)
    bufs = dict(dup_grad=torch.empty(max(st.n_dups, 1), 12, **f32), pose_partial=torch.empty(B, NB, 16, **f32),
                d_means=e(means_c), d_scales=e(scales_c), d_rots=e(rots_c), d_opac=e(opac_c), d_shs=e(shs_c), d_cols=e(cols_c),
                d_view=torch.empty(B, 16, **f32), d_raw=e(raw_c))
    gin = L.SpfRasterGradIn(_ptr(bufs["dup_grad"]), _ptr(bufs["pose_partial"]), _ptr(bufs["d_means"]),
                            _ptr(bufs["d_scales"]), _ptr(bufs["d_rots"]), _ptr(bufs["d_opac"]), _ptr(bufs["d_shs"]),
                            _ptr(bufs["d_cols"]), _ptr(bufs["d_view"]), None, _ptr(bufs["d_raw"]))
    cout = L.SpfRasterOut(_ptr(color), _ptr(depth), _ptr(alpha))
    stream = _stream(dev)
    cur = torch.cuda.current_stream(dev)
    names = [("f", i, n) for i, n in enumerate(FWD_STAGES)] + [("b", i, n) for i, n in enumerate(BWD_STAGES)]
    acc = {n: 0.0 for _, _, n in names}
    for it in range(iters + 2):
        evs = []
        for kind, i, n in names:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cur)
            if kind == "f":
                rc = lib.spf_raster_forward_stages(C.byref(st.desc), C.byref(st.cin), C.byref(st.cstate),
                                                   C.byref(cout), 1 << i, stream)
            else:
                rc = lib.spf_raster_backward_stages(C.byref(st.desc), C.byref(st.cin), C.byref(st.cstate),
                                                    C.byref(gout), C.byref(gin), 1 << i, stream)
            L.check(rc, n)
            e1.record(cur)
            evs.append((n, e0, e1))
        torch.cuda.synchronize(dev)
        if it >= 2:
            for n, e0, e1 in evs:
                acc[n] += e0.elapsed_time(e1)
    out = {n: acc[n] / iters for n in acc}
    out["_n_dups"] = st.n_dups
    if "pair_count" in st.tensors:
        pc = st.tensors["pair_count"].view(-1, 8)
        out["_pair_cap"] = int(st.desc.pair_capacity)
        out["_pair_need"] = int(st.tensors["control"][2])
        out["_fallback_tiles"] = int((pc < 0).any(dim=1).sum())
        out["_tiles"] = int(pc.shape[0])
    return out
