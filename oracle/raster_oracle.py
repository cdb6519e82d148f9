"""CPU oracle for the Gaussian-splatting hot path (TEST INFRASTRUCTURE ONLY).

This file is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product path (``spfsplatv2_b200``) must
never route through it.

PARITY UNPINNED.  The arithmetic of the reference's renderer lives in the
third-party CUDA package ``diff_gauss_pose``
(``git+https://github.com/slothfulxtx/diff-gaussian-rasterization.git@pose``,
``/root/reference/requirements.txt:87``; branch name only, no commit), which is
neither vendored under ``/root/reference`` nor installed.  The reference holds
no tests, golden vectors or fixtures for this path.  This file
restates the published 3DGS tile-rasterisation algorithm (EWA projection,
16x16 tile binning keyed by (tile, depth-bits), front-to-back alpha blending)
in plain PyTorch fp32 and anchors on the reference's own call sites:

* ``/root/reference/src/model/decoder/cuda_splatting.py:105-138`` -- the
  settings record and the forward keyword arguments (viewmatrix/projmatrix
  passed transposed, SH laid out ``[P, K, 3]``, opacities ``[P, 1]``).
* ``/root/reference/src/model/decoder/cuda_splatting.py:77-79`` -- SH degree
  ``isqrt(K) - 1`` (4 in the shipped configs).
* ``/root/reference/src/model/decoder/cuda_splatting.py:141-144`` -- only
  ``image`` and ``depth`` are consumed.

Every fp32 operation on the *index-affecting* path (projection, covariance,
radius, tile rectangle, depth key) is written as an explicit elementwise op in
a fixed order so that the CUDA kernels (compiled with ``-fmad=false`` on that
path, same order, IEEE sqrt/div) reproduce radii, rectangles, keys and sorted
lists bit for bit.  Colours, blending and gradients are compared within the
tolerances BASELINE.json's north_star states.

Everything is differentiable with torch autograd with respect to means,
scales, rotations, opacities, SH / colours and the view matrix; autograd of
this file is the ground truth for the hand-derived CUDA backward.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor

TILE = 16
NEAR_CULL = 0.2            # p_view.z <= 0.2 -> culled
COV_DILATION = 0.3         # px^2 added to the 2-D covariance diagonal
FOV_CLAMP = 1.3            # clamp of t.x/t.z, t.y/t.z in units of tan(fov/2)
ALPHA_MAX = 0.99
ALPHA_MIN = 1.0 / 255.0
T_STOP = 1e-4
EIG_FLOOR = 0.1
W_EPS = 1e-7

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
         -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658,
         0.3731763325901154, -0.4570457994644658, 1.445305721320277,
         -0.5900435899266435]
SH_C4 = [2.5033429417967046, -1.7701307697799304, 0.9461746957575601,
         -0.6690465435572892, 0.10578554691520431, -0.6690465435572892,
         0.47308734787878004, -1.7701307697799304, 0.6258357354491761]


@dataclass
class View:
    """One view's settings; mirrors GaussianRasterizationSettings + viewmatrix
    (cuda_splatting.py:105-138).  Matrices are in the row-vector convention the
    reference passes: ``p_view = [m, 1] @ viewmatrix``."""
    height: int
    width: int
    tanfovx: float
    tanfovy: float
    bg: Tensor                # [3]
    viewmatrix: Tensor        # [4,4]
    projmatrix: Tensor        # [4,4]
    sh_degree: int = 4
    scale_modifier: float = 1.0

    @property
    def grid(self):
        return ((self.width + TILE - 1) // TILE, (self.height + TILE - 1) // TILE)


def _f(v, like: Tensor) -> Tensor:
    return torch.tensor(float(v), dtype=like.dtype, device=like.device)


def sh_basis(deg: int, d: Tensor) -> Tensor:
    """Real SH basis, degrees 0..4, [P,3] unit dirs -> [P,(deg+1)^2]
    (constants/ordering of the 3DGS lineage)."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    out = [torch.full_like(x, SH_C0)]
    if deg >= 1:
        out += [-SH_C1 * y, SH_C1 * z, -SH_C1 * x]
    if deg >= 2:
        xx, yy, zz = x * x, y * y, z * z
        xy, yz, xz = x * y, y * z, x * z
        out += [SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2.0 * zz - xx - yy),
                SH_C2[3] * xz, SH_C2[4] * (xx - yy)]
    if deg >= 3:
        out += [SH_C3[0] * y * (3.0 * xx - yy), SH_C3[1] * xy * z,
                SH_C3[2] * y * (4.0 * zz - xx - yy),
                SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy),
                SH_C3[4] * x * (4.0 * zz - xx - yy), SH_C3[5] * z * (xx - yy),
                SH_C3[6] * x * (xx - 3.0 * yy)]
    if deg >= 4:
        out += [SH_C4[0] * xy * (xx - yy), SH_C4[1] * yz * (3.0 * xx - yy),
                SH_C4[2] * xy * (7.0 * zz - 1.0), SH_C4[3] * yz * (7.0 * zz - 3.0),
                SH_C4[4] * (zz * (35.0 * zz - 30.0) + 3.0),
                SH_C4[5] * xz * (7.0 * zz - 3.0),
                SH_C4[6] * (xx - yy) * (7.0 * zz - 1.0),
                SH_C4[7] * xz * (xx - 3.0 * yy),
                SH_C4[8] * (xx * (xx - 3.0 * yy) - yy * (3.0 * xx - yy))]
    return torch.stack(out, dim=-1)


def preprocess(means: Tensor, scales: Tensor, quats: Tensor, opacities: Tensor,
               shs: Optional[Tensor], colors: Optional[Tensor], view: View,
               quat_order: str = "wxyz", enable_cov_grad: bool = True,
               enable_sh_grad: bool = True) -> dict:
    """Per-Gaussian projection.  fp32, fixed op
    order.  Returns differentiable xy/depth/conic/rgb and integer radius/rect.

    ``enable_cov_grad=False`` / ``enable_sh_grad=False`` (the two switches of
    GaussianRasterizationSettings, cuda_splatting.py:117-118; both True in
    every shipped config, splatting_cuda.yaml:4-5) detach, respectively, the
    camera-dependent factor T = W J of the 2-D covariance (no gradient from the
    covariance to the mean and the pose; scales and rotations still receive
    theirs through Sigma) and the view direction of the SH colour (no gradient
    from the colour to the mean and the pose; the SH coefficients still receive
    theirs).  Forward values are unchanged."""
    assert means.dtype in (torch.float32, torch.float64)   # float64 only as a truth check for the gradient tests
    V, Pm = view.viewmatrix, view.projmatrix
    mx, my, mz = means[:, 0], means[:, 1], means[:, 2]

    def tp(j):  # p_view_j = ((V0j*mx + V1j*my) + V2j*mz) + V3j
        return ((V[0, j] * mx + V[1, j] * my) + V[2, j] * mz) + V[3, j]
    tx, ty, tz = tp(0), tp(1), tp(2)

    def ph(j):
        return ((tx * Pm[0, j] + ty * Pm[1, j]) + tz * Pm[2, j]) + Pm[3, j]
    hx, hy, hw = ph(0), ph(1), ph(3)
    p_w = 1.0 / (hw + W_EPS)
    ndcx, ndcy = hx * p_w, hy * p_w
    # NB: `python_float / tensor` is reciprocal()*float in torch; keep W,H as fp32 tensors so
    # that every division below is a true IEEE fp32 division.
    Wf, Hf = _f(view.width, means), _f(view.height, means)
    px = ((ndcx + 1.0) * Wf - 1.0) * 0.5
    py = ((ndcy + 1.0) * Hf - 1.0) * 0.5

    # 3-D covariance  Sigma = (R diag(s)) (R diag(s))^T
    if quat_order == "wxyz":
        r, x, y, z = quats[:, 0], quats[:, 1], quats[:, 2], quats[:, 3]
    else:
        x, y, z, r = quats[:, 0], quats[:, 1], quats[:, 2], quats[:, 3]
    R = [[1.0 - 2.0 * (y * y + z * z), 2.0 * (x * y - r * z), 2.0 * (x * z + r * y)],
         [2.0 * (x * y + r * z), 1.0 - 2.0 * (x * x + z * z), 2.0 * (y * z - r * x)],
         [2.0 * (x * z - r * y), 2.0 * (y * z + r * x), 1.0 - 2.0 * (x * x + y * y)]]
    s = [view.scale_modifier * scales[:, j] for j in range(3)]
    L = [[R[i][j] * s[j] for j in range(3)] for i in range(3)]
    Sg = [[(L[i][0] * L[j][0] + L[i][1] * L[j][1]) + L[i][2] * L[j][2]
           for j in range(3)] for i in range(3)]

    # 2-D covariance (EWA)
    tanx, tany = _f(view.tanfovx, means), _f(view.tanfovy, means)
    fx = Wf / (2.0 * tanx)
    fy = Hf / (2.0 * tany)
    limx, limy = FOV_CLAMP * tanx, FOV_CLAMP * tany
    txc = torch.minimum(limx, torch.maximum(-limx, tx / tz)) * tz
    tyc = torch.minimum(limy, torch.maximum(-limy, ty / tz)) * tz
    tz2 = tz * tz
    J00 = fx / tz
    J02 = -(fx * txc) / tz2
    J11 = fy / tz
    J12 = -(fy * tyc) / tz2
    T0 = [J00 * V[i, 0] + J02 * V[i, 2] for i in range(3)]
    T1 = [J11 * V[i, 1] + J12 * V[i, 2] for i in range(3)]
    if not enable_cov_grad:
        T0 = [x.detach() for x in T0]
        T1 = [x.detach() for x in T1]
    U0 = [(T0[0] * Sg[0][j] + T0[1] * Sg[1][j]) + T0[2] * Sg[2][j] for j in range(3)]
    U1 = [(T1[0] * Sg[0][j] + T1[1] * Sg[1][j]) + T1[2] * Sg[2][j] for j in range(3)]
    a = ((U0[0] * T0[0] + U0[1] * T0[1]) + U0[2] * T0[2]) + COV_DILATION
    b = (U0[0] * T1[0] + U0[1] * T1[1]) + U0[2] * T1[2]
    c = ((U1[0] * T1[0] + U1[1] * T1[1]) + U1[2] * T1[2]) + COV_DILATION
    det = a * c - b * b
    det_inv = 1.0 / det
    conic = torch.stack([c * det_inv, -b * det_inv, a * det_inv], dim=-1)
    mid = 0.5 * (a + c)
    root = torch.sqrt(torch.clamp_min(mid * mid - det, EIG_FLOOR))
    lam = torch.maximum(mid + root, mid - root)
    radius_f = torch.ceil(3.0 * torch.sqrt(lam))

    gx, gy = view.grid
    with torch.no_grad():
        def tile_lo(p, r, g):
            return torch.clamp((p - r) / float(TILE), 0.0, float(g)).trunc().to(torch.int32)

        def tile_hi(p, r, g):
            return torch.clamp(((p + r) + float(TILE - 1)) / float(TILE), 0.0, float(g)).trunc().to(torch.int32)
        rect = torch.stack([tile_lo(px, radius_f, gx), tile_lo(py, radius_f, gy),
                            tile_hi(px, radius_f, gx), tile_hi(py, radius_f, gy)], dim=-1)
        area = (rect[:, 2] - rect[:, 0]) * (rect[:, 3] - rect[:, 1])
        visible = (tz > NEAR_CULL) & (det != 0) & (area > 0)
        visible &= torch.isfinite(radius_f) & torch.isfinite(px) & torch.isfinite(py)
        radius = torch.where(visible, radius_f, torch.zeros_like(radius_f)).to(torch.int32)
        tiles_touched = torch.where(visible, area, torch.zeros_like(area))
        rect = torch.where(visible[:, None], rect, torch.zeros_like(rect))

    # colour
    if shs is not None:
        A, tau = V[:3, :3], V[3, :3]
        campos = torch.stack([-((tau[0] * A[i, 0] + tau[1] * A[i, 1]) + tau[2] * A[i, 2])
                              for i in range(3)])
        d = means - campos[None, :]
        dn = d / torch.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])[:, None]
        if not enable_sh_grad:
            dn = dn.detach()
        basis = sh_basis(view.sh_degree, dn)                     # [P,K]
        K = basis.shape[1]
        rgb = (basis[:, :, None] * shs[:, :K, :]).sum(dim=1) + 0.5
        rgb = torch.clamp_min(rgb, 0.0)
    else:
        rgb = colors

    return dict(xy=torch.stack([px, py], dim=-1), depth=tz, conic=conic,
                opacity=opacities.reshape(-1), rgb=rgb, radius=radius, rect=rect,
                tiles_touched=tiles_touched, visible=visible,
                cov2d=torch.stack([a, b, c], dim=-1))


def bin_and_sort(pre: dict, view: View):
    """Duplicate-with-keys + stable sort + tile ranges (SURVEY.md App. B
    'Binning').  Returns (keys int64 [N], point_list int32 [N], ranges int32
    [T,2]); key = (tile << 32) | float_bits(depth)."""
    gx, gy = view.grid
    rect = pre["rect"].to(torch.int64)
    vis = torch.nonzero(pre["visible"]).flatten()
    depth_bits = pre["depth"].detach().float().contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    keys, vals = [], []
    if vis.numel():
        r = rect[vis]
        w = r[:, 2] - r[:, 0]
        n = w * (r[:, 3] - r[:, 1])
        owner = torch.repeat_interleave(torch.arange(vis.numel()), n)
        start = torch.cumsum(n, 0) - n
        local = torch.arange(int(n.sum())) - start[owner]
        ty = r[owner, 1] + local // w[owner]
        tx = r[owner, 0] + local % w[owner]
        tile = ty * gx + tx
        g = vis[owner]
        keys = (tile << 32) | depth_bits[g]
        vals = g
        order = torch.sort(keys, stable=True).indices
        keys, vals = keys[order], vals[order].to(torch.int32)
    else:
        keys = torch.zeros(0, dtype=torch.int64)
        vals = torch.zeros(0, dtype=torch.int32)
    T = gx * gy
    ranges = torch.zeros(T, 2, dtype=torch.int32)
    if keys.numel():
        tile_of = (keys >> 32)
        counts = torch.bincount(tile_of, minlength=T)
        ends = torch.cumsum(counts, 0)
        starts = ends - counts
        nz = counts > 0
        ranges[nz, 0] = starts[nz].to(torch.int32)
        ranges[nz, 1] = ends[nz].to(torch.int32)
    return keys, vals, ranges


def blend(pre: dict, point_list: Tensor, ranges: Tensor, view: View,
          clamp_straight_through: bool = True):
    """Per-16x16-tile front-to-back alpha blend (SURVEY.md App. B 'Blend
    forward').  Returns color [3,H,W], depth [1,H,W], alpha [1,H,W],
    final_T [H,W], n_contrib int32 [H,W].

    ``clamp_straight_through``: the upstream lineage's backward ignores the
    min(0.99, .) saturation (gradient passes as if unclamped); True mirrors
    that, False gives the exact derivative of the forward."""
    H, W = view.height, view.width
    gx, gy = view.grid
    xy, conic, opac, rgb, depth = pre["xy"], pre["conic"], pre["opacity"], pre["rgb"], pre["depth"]
    dt = xy.dtype
    color = torch.zeros(3, H, W, dtype=dt)
    dimg = torch.zeros(1, H, W, dtype=dt)
    final_T = torch.ones(H, W, dtype=dt)
    n_contrib = torch.zeros(H, W, dtype=torch.int32)
    col_tiles, dep_tiles, T_tiles = {}, {}, {}
    for t in range(gx * gy):
        s, e = int(ranges[t, 0]), int(ranges[t, 1])
        if e <= s:
            continue
        ty_, tx_ = divmod(t, gx)
        y0, x0 = ty_ * TILE, tx_ * TILE
        y1, x1 = min(y0 + TILE, H), min(x0 + TILE, W)
        ys = torch.arange(y0, y1, dtype=dt)
        xs = torch.arange(x0, x1, dtype=dt)
        pyy, pxx = torch.meshgrid(ys, xs, indexing="ij")
        pxx, pyy = pxx.reshape(-1), pyy.reshape(-1)
        ids = point_list[s:e].long()
        dx = xy[ids, 0][:, None] - pxx[None, :]
        dy = xy[ids, 1][:, None] - pyy[None, :]
        cn = conic[ids]
        power = -0.5 * (cn[:, 0:1] * dx * dx + cn[:, 2:3] * dy * dy) - cn[:, 1:2] * dx * dy
        raw = opac[ids][:, None] * torch.exp(power)
        if clamp_straight_through:
            alpha = raw + (torch.clamp_max(raw, ALPHA_MAX) - raw).detach()
        else:
            alpha = torch.clamp_max(raw, ALPHA_MAX)
        valid = (power <= 0) & (alpha >= ALPHA_MIN)
        a = torch.where(valid, alpha, torch.zeros_like(alpha))
        one_m = 1.0 - a
        Tincl = torch.cumprod(one_m, dim=0)
        Texcl = torch.cat([torch.ones_like(Tincl[:1]), Tincl[:-1]], dim=0)
        stop = valid & (Tincl.detach() < T_STOP)
        stopped = torch.cumsum(stop.to(torch.int32), dim=0) > 0
        live = valid & ~stopped
        w = torch.where(live, a * Texcl, torch.zeros_like(a))
        c_t = (w[:, :, None] * rgb[ids][:, None, :]).sum(0)            # [px,3]
        d_t = (w * depth[ids][:, None]).sum(0)
        fT = torch.where(live, one_m, torch.ones_like(one_m)).prod(dim=0)
        idx = torch.arange(1, e - s + 1, dtype=torch.int32)[:, None]
        nc = torch.where(live, idx, torch.zeros_like(idx)).max(dim=0).values
        hh, ww = y1 - y0, x1 - x0
        col_tiles[t] = (y0, y1, x0, x1, c_t.t().reshape(3, hh, ww))
        dep_tiles[t] = d_t.reshape(hh, ww)
        T_tiles[t] = fT.reshape(hh, ww)
        n_contrib[y0:y1, x0:x1] = nc.reshape(hh, ww)
    # assemble without in-place writes into leaf-less zeros (keeps autograd simple)
    color = color.clone()
    for t, (y0, y1, x0, x1, c_t) in col_tiles.items():
        color[:, y0:y1, x0:x1] = c_t
        dimg[0, y0:y1, x0:x1] = dep_tiles[t]
        final_T[y0:y1, x0:x1] = T_tiles[t]
    color = color + final_T[None] * view.bg.to(dt).reshape(3, 1, 1)
    alpha_img = (1.0 - final_T)[None]
    return color, dimg, alpha_img, final_T, n_contrib


def render(means, scales, quats, opacities, shs, colors, view: View,
           quat_order: str = "wxyz", clamp_straight_through: bool = True,
           enable_cov_grad: bool = True, enable_sh_grad: bool = True) -> dict:
    """Full forward for one view.  Differentiable outputs: color, depth, alpha."""
    pre = preprocess(means, scales, quats, opacities, shs, colors, view, quat_order,
                     enable_cov_grad, enable_sh_grad)
    keys, point_list, ranges = bin_and_sort(pre, view)
    color, depth, alpha, final_T, n_contrib = blend(pre, point_list, ranges, view,
                                                    clamp_straight_through)
    return dict(color=color, depth=depth, alpha=alpha, final_T=final_T,
                n_contrib=n_contrib, keys=keys, point_list=point_list,
                ranges=ranges, pre=pre)


def compute_psnr(gt: Tensor, pred: Tensor) -> Tensor:
    """PSNR as the reference defines it (src/evaluation/metrics.py:12-19)."""
    gt = gt.clip(0, 1)
    pred = pred.clip(0, 1)
    mse = ((gt - pred) ** 2).flatten(1).mean(dim=1)
    return -10 * mse.log10()
