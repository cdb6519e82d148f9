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
_capacity_hint: dict = {}
GROWTH = 1.25
_host_counters: dict = {}     # device index -> (pinned int32[2] tensor, numpy view)
_ticket = [0]
# Pair-log sizing (records per warp): per-shape high-water mark of control[2], fed back asynchronously after each
# backward (pinned copy + event, polled -- never waited on -- by the next forward of that shape).
_pair_cap_hint: dict = {}
_pair_stat: dict = {}         # key -> [[pinned int32[1], event, age in forwards], ...]
_pin_pool: list = []          # recycled (pinned int32[1], event) pairs: cudaHostAlloc per step would cost ~100 us
PAIR_CAP_DEFAULT = 512
PAIR_LOG_BUDGET = 4 << 30     # bytes; above this the capacity is clipped and the densest tiles fall back to recomputation


def _counters(dev):
    hc = _host_counters.get(dev.index)
    if hc is None:
        t = torch.zeros(2, dtype=torch.int32).pin_memory()
        hc = (t, t.numpy())
        _host_counters[dev.index] = hc
    return hc


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


class _State:
    """Forward intermediates kept for backward / inspection (plain tensors, not autograd-tracked)."""
    __slots__ = ("desc", "cin", "cstate", "keep", "n_dups", "capacity", "tensors", "key")


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
    return max(64, min(cap, (PAIR_LOG_BUDGET // (32 * max(n_warps, 1))) // 64 * 64))


def _forward_impl(s: RasterSettings, means, scales, rots, opac, shs, colors, viewmat, projmat, tanfov, bg,
                  pre_scale, pair_log: bool = False):
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
    if use_sh == (colors is not None):
        raise Exception("Please provide excatly one of either SHs or precomputed colors!")
    K = 0
    if use_sh:
        K = shs.shape[-1] if s.sh_layout_ck else shs.shape[-2]

    f32 = dict(dtype=torch.float32, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    key = (dev.index, S, v, P, H, W)
    cap = max(int(_capacity_hint.get(key, 2 * B * P)), 1024)

    pair_cap = _pair_capacity(key, B * T * 8) if (pair_log and s.pair_log) else 0
    desc = L.SpfRasterDesc(S, v, P, H, W, s.sh_degree, s.flags(), float(s.scale_modifier), cap, 0, pair_cap)
    n_ctrl = lib.spf_raster_control_ints(C.byref(desc))
    if n_ctrl < 0:
        L.check(-1, "spf_raster_control_ints")

    cin = L.SpfRasterIn(_ptr(means), _ptr(scales), _ptr(rots), _ptr(opac), _ptr(shs), _ptr(colors), K,
                        _ptr(viewmat), _ptr(projmat), _ptr(tanfov), _ptr(bg), _ptr(pre_scale))
    t = dict(
        xy=torch.empty(B, P, 2, **f32), depth=torch.empty(B, P, **f32), conic_opacity=torch.empty(B, P, 4, **f32),
        rgb=torch.empty(B, P, 3, **f32), radii=torch.empty(B, P, **i32), tiles_touched=torch.empty(B, P, **i32),
        dup_offset=torch.empty(B, P, **i32), control=torch.empty(n_ctrl, **i32),
        tile_ranges=torch.empty(B * T, 2, **i32), final_T=torch.empty(B, H, W, **f32),
        n_contrib=torch.empty(B, H, W, **i32), accum=torch.empty(B, H, W, 4, **f32))
    if pair_cap > 0:
        t["pair_log"] = torch.empty(B * T * 8, pair_cap, 8, **f32)
        t["pair_count"] = torch.empty(B * T * 8, **i32)
    color = torch.empty(B, 3, H, W, **f32)
    depth = torch.empty(B, 1, H, W, **f32)
    alpha = torch.empty(B, 1, H, W, **f32) if s.want_alpha else None
    host_t, host_np = _counters(dev)
    stream = _stream(dev)
    while True:
        _ticket[0] = (_ticket[0] % 0x3fffffff) + 1
        ticket = _ticket[0]
        desc.dup_capacity = cap
        desc.ticket = ticket
        t["bucket"] = torch.empty(cap, dtype=torch.int64, device=dev)
        t["slab"] = torch.empty(cap, 12, **f32)
        t["cullbox"] = torch.empty(cap, 4, **f32)
        cstate = L.SpfRasterState(*[_ptr(t[k]) for k in ("xy", "depth", "conic_opacity", "rgb", "radii",
                                                        "tiles_touched", "dup_offset", "control", "bucket",
                                                        "slab", "cullbox", "tile_ranges", "final_T", "n_contrib", "accum")],
                                  _ptr(t.get("pair_log")), _ptr(t.get("pair_count")), _ptr(host_t))
        cout = L.SpfRasterOut(_ptr(color), _ptr(depth), _ptr(alpha))
        L.check(lib.spf_raster_forward(C.byref(desc), C.byref(cin), C.byref(cstate), C.byref(cout), stream),
                "spf_raster_forward")
        # The scan kernel stores {N, ticket} into mapped pinned memory; poll for our ticket.  All kernels of
        # this forward are already queued, so the GPU keeps running emit / sort / blend meanwhile.
        spins = 0
        while host_np[1] != ticket:
            spins += 1
            if spins > 2_000_000 and (spins & 0xfffff) == 0:
                torch.cuda.current_stream(dev).synchronize()   # surfaces a launch failure instead of hanging
                if host_np[1] != ticket:
                    raise RuntimeError("spf_raster_forward: duplicate count never arrived (kernel failure?)")
        n_dups = int(host_np[0])
        if n_dups <= cap:
            break
        cap = int(n_dups * 1.05) + 1024
    _capacity_hint[key] = max(int(n_dups * GROWTH) + 1024, 1024)

    st = _State()
    st.desc, st.cin, st.cstate, st.n_dups, st.capacity, st.tensors = desc, cin, cstate, n_dups, cap, t
    st.key = key
    st.keep = (means, scales, rots, opac, shs, colors, viewmat, projmat, tanfov, bg, pre_scale)
    return color, depth, alpha, t["radii"], st


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, settings: RasterSettings, means, scales, rots, opac, shs, colors, viewmat, projmat, tanfov,
                bg, pre_scale, means2d):
        args = [None if a is None else _f32c(a.detach()) for a in
                (means, scales, rots, opac, shs, colors, viewmat, projmat, tanfov, bg, pre_scale)]
        color, depth, alpha, radii, st = _forward_impl(settings, *args, pair_log=any(ctx.needs_input_grad))
        ctx.settings = settings
        ctx.st = st
        ctx.shapes = (means.shape, scales.shape, rots.shape, opac.shape,
                      None if shs is None else shs.shape, None if colors is None else colors.shape,
                      viewmat.shape, None if means2d is None else means2d.shape)
        ctx.mark_non_differentiable(radii)
        if alpha is None:
            alpha = torch.empty(0, device=color.device)
            ctx.mark_non_differentiable(alpha)
        return color, depth, alpha, radii

    @staticmethod
    def backward(ctx, g_color, g_depth, g_alpha, _g_radii):
        s: RasterSettings = ctx.settings
        st: _State = ctx.st
        lib = L.lib()
        means, scales, rots, opac, shs, colors, viewmat, projmat, tanfov, bg, pre_scale = st.keep
        dev = means.device
        S, P = means.shape[0], means.shape[1]
        B = S * s.views_per_scene
        NB = (P + PROJ_THREADS - 1) // PROJ_THREADS
        f32 = dict(dtype=torch.float32, device=dev)
        gc = None if g_color is None else _f32c(g_color)
        gd = None if g_depth is None else _f32c(g_depth)
        ga = None if (g_alpha is None or not s.want_alpha) else _f32c(g_alpha)
        gout = L.SpfRasterGradOut(_ptr(gc), _ptr(gd), _ptr(ga))
        dup_grad = torch.empty(max(st.n_dups, 1), 12, **f32)
        pose_partial = torch.empty(B, NB, 16, **f32)
        d_means = torch.empty_like(means)
        d_scales = torch.empty_like(scales)
        d_rots = torch.empty_like(rots)
        d_opac = torch.empty_like(opac)
        d_shs = torch.empty_like(shs) if shs is not None else None
        d_cols = torch.empty_like(colors) if colors is not None else None
        d_view = torch.empty(B, 16, **f32)
        d_m2d = torch.empty(B, P, 3, **f32) if (s.want_means2d_grad and ctx.shapes[7] is not None) else None
        gin = L.SpfRasterGradIn(_ptr(dup_grad), _ptr(pose_partial), _ptr(d_means), _ptr(d_scales), _ptr(d_rots),
                                _ptr(d_opac), _ptr(d_shs), _ptr(d_cols), _ptr(d_view), _ptr(d_m2d))
        L.check(lib.spf_raster_backward(C.byref(st.desc), C.byref(st.cin), C.byref(st.cstate), C.byref(gout),
                                        C.byref(gin), _stream(dev)), "spf_raster_backward")
        if st.desc.pair_capacity > 0 and len(_pair_stat.setdefault(st.key, [])) < 4:
            # feed the largest per-warp pair count back to the sizing of the next forward (no wait)
            pin, ev = _pin_pool.pop() if _pin_pool else (torch.empty(1, dtype=torch.int32).pin_memory(), torch.cuda.Event())
            pin.copy_(st.tensors["control"][2:3], non_blocking=True)
            ev.record(torch.cuda.current_stream(dev))
            _pair_stat[st.key].append([pin, ev, 0])
        sh = ctx.shapes
        return (None, d_means.view(sh[0]), d_scales.view(sh[1]), d_rots.view(sh[2]), d_opac.view(sh[3]),
                None if d_shs is None else d_shs.view(sh[4]), None if d_cols is None else d_cols.view(sh[5]),
                d_view.view(sh[6]), None, None, None, None,
                None if d_m2d is None else d_m2d.view(sh[7]))


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


def forward_with_state(settings: RasterSettings, means, scales, rotations, opacities, shs, colors, viewmatrix,
                       projmatrix, tanfov, bg, pre_scale=None, pair_log: bool = False):
    """No-autograd forward that also returns the intermediate state (for parity tests / profiling)."""
    args = [None if a is None else _f32c(a.detach()) for a in
            (means, scales, rotations, opacities, shs, colors, viewmatrix, projmatrix, tanfov, bg, pre_scale)]
    return _forward_impl(settings, *args, pair_log=pair_log)


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
                   projmatrix, tanfov, bg, pre_scale, g_color, g_depth, iters: int = 10) -> dict:
    """Per-kernel device times (ms, mean over ``iters``) measured with CUDA events around each stage of
    the forward and backward launch sequences (spf_raster_{forward,backward}_stages), on the current
    stream.  Used by bench.py for the roofline of the dominant kernel."""
    lib = L.lib()
    color, depth, alpha, radii, st = forward_with_state(settings, means, scales, rotations, opacities, shs, colors,
                                                        viewmatrix, projmatrix, tanfov, bg, pre_scale, pair_log=True)
    means_c, scales_c, rots_c, opac_c, shs_c, cols_c = st.keep[:6]
    dev = means_c.device
    S, P = means_c.shape[0], means_c.shape[1]
    B = S * settings.views_per_scene
    NB = (P + PROJ_THREADS - 1) // PROJ_THREADS
    f32 = dict(dtype=torch.float32, device=dev)
    gc, gd = _f32c(g_color), (None if g_depth is None else _f32c(g_depth))
    gout = L.SpfRasterGradOut(_ptr(gc), _ptr(gd), None)
    bufs = dict(dup_grad=torch.empty(max(st.n_dups, 1), 12, **f32), pose_partial=torch.empty(B, NB, 16, **f32),
                d_means=torch.empty_like(means_c), d_scales=torch.empty_like(scales_c), d_rots=torch.empty_like(rots_c),
                d_opac=torch.empty_like(opac_c), d_shs=None if shs_c is None else torch.empty_like(shs_c),
                d_cols=None if cols_c is None else torch.empty_like(cols_c), d_view=torch.empty(B, 16, **f32))
    gin = L.SpfRasterGradIn(_ptr(bufs["dup_grad"]), _ptr(bufs["pose_partial"]), _ptr(bufs["d_means"]),
                            _ptr(bufs["d_scales"]), _ptr(bufs["d_rots"]), _ptr(bufs["d_opac"]), _ptr(bufs["d_shs"]),
                            _ptr(bufs["d_cols"]), _ptr(bufs["d_view"]), None)
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
