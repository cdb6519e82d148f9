"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Bars (BASELINE.json north_star): tile/sort indices bit-exact; images |dPSNR| < 1e-3 dB;
gradients within 1e-4 relative."""
import math

import pytest
import torch

from oracle import raster_oracle as O
from spfsplatv2_b200.synthetic import make_batch, make_scene
from tests.util import oracle_views, rel_err

pytestmark = pytest.mark.gpu

GRAD_TOL = 1e-4
PSNR_TOL = 1e-3


def _dev():
    return torch.device("cuda:0")


def _decoder(bg=(0.0, 0.0, 0.0), scale_invariant=True):
    from spfsplatv2_b200.decoder import DecoderSplattingCUDA, DecoderSplattingCUDACfg
    return DecoderSplattingCUDA(DecoderSplattingCUDACfg("splatting_cuda", list(bg), scale_invariant, True, True)).to(_dev())


def _gaussians(sc, requires_grad=False):
    from spfsplatv2_b200.decoder import Gaussians
    d = _dev()
    t = {k: getattr(sc, k).to(d) for k in ("means", "rotations", "scales", "harmonics", "opacities")}
    if requires_grad:
        for x in t.values():
            x.requires_grad_()
    return Gaussians(t["means"], sc.covariances.to(d), t["rotations"], t["scales"], t["harmonics"], t["opacities"]), t


def _state_forward(sc, no_tma=False, bg=(0.0, 0.0, 0.0)):
    """Raw batched forward returning the intermediate state (for index parity)."""
    from spfsplatv2_b200.camera import camera_setup
    from spfsplatv2_b200.rasterizer import RasterSettings, forward_with_state
    d = _dev()
    b, v = sc.extrinsics.shape[:2]
    h, w = sc.image_shape
    # camera matrices on the CPU (exactly the oracle's inputs: a GPU matrix inverse differs by ulps, and the
    # index parity below is bit-exact), then moved to the device
    view, proj, tanfov, scale = [x.to(d) for x in camera_setup(
        sc.extrinsics.reshape(b * v, 4, 4), sc.intrinsics.reshape(b * v, 3, 3), sc.near.reshape(-1),
        sc.far.reshape(-1), True)]
    K = sc.harmonics.shape[-1]
    s = RasterSettings(h, w, math.isqrt(K) - 1, 1.0, v, sh_layout_ck=True, want_alpha=True, no_tma=no_tma)
    bgt = torch.tensor(bg, dtype=torch.float32, device=d).expand(b * v, 3)
    return forward_with_state(s, sc.means.to(d), sc.scales.to(d), sc.rotations.to(d), sc.opacities.to(d),
                              sc.harmonics.to(d), None, view, proj, tanfov, bgt, scale)


@pytest.mark.parametrize("regime,h,w,grid,v_cxt", [
    ("init", 64, 64, (32, 32), 1),        # BASELINE config 1: 1k Gaussians -> 64x64
    ("trained", 64, 48, (32, 32), 1),     # non-square, partial edge tiles
    ("trained", 100, 72, (40, 40), 2),    # sizes not multiples of 16
    ("init", 256, 256, None, 1),          # headline "65k" scene
])
def test_indices_bit_exact(regime, h, w, grid, v_cxt):
    from spfsplatv2_b200.rasterizer import unpack_sorted
    sc = make_scene(seed=3, v_cxt=v_cxt, h=h, w=w, grid=grid, regime=regime, n_target=2)
    color, depth, alpha, radii, st = _state_forward(sc)
    ref, _ = oracle_views(sc)
    pl, keys = unpack_sorted(st)
    pl, keys = pl.cpu(), keys.cpu()
    T = st.tensors["tile_ranges"].shape[0] // 2
    ranges = st.tensors["tile_ranges"].cpu().view(2, T, 2)
    tiles = st.tensors["tiles_touched"].cpu()
    off = 0
    for i, r in enumerate(ref):
        assert torch.equal(radii[i].cpu(), r["pre"]["radius"]), f"view {i}: radii differ"
        assert torch.equal(tiles[i], r["pre"]["tiles_touched"].to(torch.int32)), f"view {i}: tiles_touched differ"
        n = r["keys"].numel()
        assert torch.equal(keys[off:off + n], r["keys"]), f"view {i}: sorted keys differ"
        assert torch.equal(pl[off:off + n], r["point_list"]), f"view {i}: sorted point list differs"
        rg = ranges[i].clone()
        nz = rg[:, 1] > rg[:, 0]
        rg[nz] -= off
        assert torch.equal(rg, r["ranges"]), f"view {i}: tile ranges differ"
        assert torch.equal(st.tensors["n_contrib"][i].cpu(), r["n_contrib"]), f"view {i}: n_contrib differs"
        off += n
    assert off == st.n_dups


def test_sort_degenerate_and_long_lists_bit_exact():
    """Per-tile bucket sort edge cases: (a) hundreds of Gaussians with EXACTLY the same depth in one tile (one bucket
    overflows -> bitonic fallback; ties must come out in Gaussian-id order), (b) a tile list far longer than the
    shared-memory capacity chosen from the average (in-place global sort)."""
    from spfsplatv2_b200.rasterizer import unpack_sorted
    sc = make_scene(seed=43, v_cxt=1, h=64, w=64, grid=(48, 48), regime="trained", n_target=1)
    sc.means[0, 100:400] = sc.means[0, 100]          # 300 coincident Gaussians: identical depth bits
    sc.means[0, 1000:2200, :2] *= 0.02                # 1200 Gaussians piled into the central tiles (long lists)
    color, depth, alpha, radii, st = _state_forward(sc)
    ref, _ = oracle_views(sc)
    pl, keys = unpack_sorted(st)
    r = ref[0]
    n = r["keys"].numel()
    assert st.n_dups == n
    assert torch.equal(keys.cpu(), r["keys"]) and torch.equal(pl.cpu(), r["point_list"])
    lens = (st.tensors["tile_ranges"][:, 1] - st.tensors["tile_ranges"][:, 0]).cpu()
    assert int(lens.max()) > 1024                     # the long-list path was exercised
    assert (color[0].cpu() - r["color"]).abs().max().item() < 3e-5


@pytest.mark.parametrize("regime,h,w,grid,bg", [
    ("init", 64, 64, (32, 32), (0.0, 0.0, 0.0)),
    ("trained", 64, 48, (32, 32), (0.3, 0.5, 0.7)),
    ("trained", 128, 128, (64, 64), (0.0, 0.0, 0.0)),
])
def test_image_parity(regime, h, w, grid, bg):
    sc = make_scene(seed=5, v_cxt=1, h=h, w=w, grid=grid, regime=regime, n_target=2)
    color, depth, alpha, radii, st = _state_forward(sc, bg=bg)
    ref, _ = oracle_views(sc, bg=bg)
    gt, _ = oracle_views(make_scene(seed=6, v_cxt=1, h=h, w=w, grid=grid, regime=regime, n_target=2), bg=bg)
    for i, r in enumerate(ref):
        c = color[i].cpu()
        assert (c - r["color"]).abs().max().item() < 2e-5
        assert (depth[i].cpu() - r["depth"]).abs().max().item() < 2e-4
        assert (alpha[i].cpu() - r["alpha"]).abs().max().item() < 2e-5
        pseudo_gt = gt[i]["color"][None]
        dpsnr = (O.compute_psnr(pseudo_gt, c[None]) - O.compute_psnr(pseudo_gt, r["color"][None])).abs().item()
        assert dpsnr < PSNR_TOL, dpsnr


def test_tma_and_plain_staging_agree():
    sc = make_scene(seed=7, v_cxt=1, h=128, w=128, grid=(64, 64), regime="trained", n_target=1)
    a = _state_forward(sc, no_tma=False)
    b = _state_forward(sc, no_tma=True)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert torch.equal(a[4].tensors["n_contrib"], b[4].tensors["n_contrib"])


def _cuda_render_identical_inputs(sc, bg, requires_grad=True):
    """The CUDA rasterizer on bit-identical inputs to the oracle's: camera matrices are computed on the CPU by the
    same torch glue (a GPU matrix inverse differs by ulps, and one alpha >= 1/255 membership flip at a single pixel
    changes a gradient by ~1e-2 absolute -- the blend is discontinuous there), then moved to the device.  Autograd
    flows back through the move to the CPU extrinsics leaf."""
    from spfsplatv2_b200.camera import camera_setup
    from spfsplatv2_b200.rasterizer import RasterSettings, rasterize_batched
    d = _dev()
    b, v = sc.extrinsics.shape[:2]
    h, w = sc.image_shape
    ext = sc.extrinsics.clone().requires_grad_(requires_grad)
    view, proj, tanfov, scale = camera_setup(ext.reshape(b * v, 4, 4), sc.intrinsics.reshape(b * v, 3, 3),
                                             sc.near.reshape(-1), sc.far.reshape(-1), True)
    t = {k: getattr(sc, k).to(d).requires_grad_(requires_grad) for k in ("means", "rotations", "scales", "harmonics", "opacities")}
    K = sc.harmonics.shape[-1]
    s = RasterSettings(h, w, math.isqrt(K) - 1, 1.0, v, sh_layout_ck=True)
    bgt = torch.tensor(bg, dtype=torch.float32, device=d).expand(b * v, 3)
    color, depth, _, _ = rasterize_batched(s, t["means"], t["scales"], t["rotations"], t["opacities"], t["harmonics"], None,
                                           view.to(d), proj.to(d), tanfov.to(d), bgt, scale.to(d))
    depth = depth * sc.near.reshape(-1).to(d)[:, None, None, None]
    return color, depth, t, ext


@pytest.mark.parametrize("regime,h,w,grid,b,v", [
    ("init", 64, 64, (32, 32), 1, 1),
    ("trained", 64, 48, (24, 24), 2, 3),   # several views per scene: gradients sum over views
    ("trained", 96, 96, (48, 48), 1, 1),
    ("trained", 128, 128, (64, 64), 1, 2),
])
def test_gradients_match_oracle_autograd(regime, h, w, grid, b, v):
    sc = make_batch(b, seed=11, v_cxt=1, h=h, w=w, grid=grid, regime=regime, n_target=v, with_cov=True)
    bg = (0.2, 0.1, 0.4)
    ref, leaves = oracle_views(sc, bg=bg, requires_grad=True)
    torch.manual_seed(0)
    wc = torch.randn(b * v, 3, h, w)
    wd = 0.05 * torch.randn(b * v, 1, h, w)
    loss = sum((r["color"] * wc[i]).sum() + (r["depth"] * sc.near.reshape(-1)[i] * wd[i]).sum() for i, r in enumerate(ref))
    loss.backward()

    color, depth, t, ext = _cuda_render_identical_inputs(sc, bg)
    l2 = (color * wc.to(_dev())).sum() + (depth * wd.to(_dev())).sum()
    l2.backward()
    assert abs(l2.item() - loss.item()) <= 1e-4 * max(1.0, abs(loss.item()))
    for name in ("means", "scales", "rotations", "opacities", "harmonics"):
        e = rel_err(t[name].grad.cpu(), leaves[name].grad)
        assert e < GRAD_TOL, f"{name}: rel err {e:.3e}"
    e = rel_err(ext.grad, leaves["extrinsics"].grad)
    assert e < GRAD_TOL, f"extrinsics (pose): rel err {e:.3e}"


@pytest.mark.parametrize("cov,sh,xyzw", [(False, True, False), (True, False, False), (False, False, False), (True, True, True)])
def test_gradient_switches_and_quaternion_order(cov, sh, xyzw):
    """GaussianRasterizationSettings.enable_cov_grad / enable_sh_grad = False (cuda_splatting.py:117-118,
    config/model/decoder/splatting_cuda.yaml:4-5) in the BACKWARD, against the oracle with the corresponding factors
    detached; and rotations given as (x,y,z,w) (SPF_FLAG_QUAT_XYZW), forward and backward."""
    from spfsplatv2_b200.camera import camera_setup
    from spfsplatv2_b200.rasterizer import RasterSettings, rasterize_batched
    d = _dev()
    b, v, h, w = 1, 2, 80, 64
    sc = make_batch(b, seed=71, v_cxt=1, h=h, w=w, grid=(40, 40), regime="trained", n_target=v, with_cov=True)
    bg = (0.2, 0.1, 0.4)
    ref, leaves = oracle_views(sc, bg=bg, requires_grad=True, enable_cov_grad=cov, enable_sh_grad=sh,
                               quat_order="xyzw" if xyzw else "wxyz")
    torch.manual_seed(1)
    wc = torch.randn(b * v, 3, h, w)
    wd = 0.05 * torch.randn(b * v, 1, h, w)
    loss = sum((r["color"] * wc[i]).sum() + (r["depth"] * wd[i]).sum() for i, r in enumerate(ref))
    loss.backward()
    ext = sc.extrinsics.clone().requires_grad_()
    view, proj, tanfov, scale = camera_setup(ext.reshape(b * v, 4, 4), sc.intrinsics.reshape(b * v, 3, 3),
                                             sc.near.reshape(-1), sc.far.reshape(-1), True)
    t = {k: getattr(sc, k).to(d).requires_grad_() for k in ("means", "rotations", "scales", "harmonics", "opacities")}
    s = RasterSettings(h, w, 4, 1.0, v, sh_layout_ck=True, enable_cov_grad=cov, enable_sh_grad=sh, quat_xyzw=xyzw)
    bgt = torch.tensor(bg, dtype=torch.float32, device=d).expand(b * v, 3)
    color, depth, _, _ = rasterize_batched(s, t["means"], t["scales"], t["rotations"], t["opacities"], t["harmonics"], None,
                                           view.to(d), proj.to(d), tanfov.to(d), bgt, scale.to(d))
    for i, r in enumerate(ref):
        assert (color[i].cpu() - r["color"]).abs().max().item() < 3e-5
    ((color * wc.to(d)).sum() + (depth * wd.to(d)).sum()).backward()
    for name in ("means", "scales", "rotations", "opacities", "harmonics"):
        e = rel_err(t[name].grad.cpu(), leaves[name].grad)
        assert e < GRAD_TOL, f"{name}: rel err {e:.3e}"
    assert rel_err(ext.grad, leaves["extrinsics"].grad) < GRAD_TOL
    if not cov or not sh:      # the switches do change the gradient: the full gradient must NOT match
        ref2, leaves2 = oracle_views(sc, bg=bg, requires_grad=True)
        sum((r["color"] * wc[i]).sum() + (r["depth"] * wd[i]).sum() for i, r in enumerate(ref2)).backward()
        assert rel_err(t["means"].grad.cpu(), leaves2["means"].grad) > 10 * GRAD_TOL


def test_decoder_gradients_end_to_end():
    """Whole public path (DecoderSplattingCUDA with the fused GPU camera-setup kernel).  The GPU matrix inverse
    differs from the CPU one by ulps, which may flip individual alpha-threshold memberships (see above), so the
    bar here is 1e-3; the 1e-4 bar is enforced on bit-identical inputs in test_gradients_match_oracle_autograd."""
    b, v, h, w = 1, 1, 96, 96
    sc = make_batch(b, seed=11, v_cxt=1, h=h, w=w, grid=(48, 48), regime="trained", n_target=v, with_cov=True)
    bg = (0.2, 0.1, 0.4)
    ref, leaves = oracle_views(sc, bg=bg, requires_grad=True)
    torch.manual_seed(0)
    wc = torch.randn(b * v, 3, h, w)
    wd = 0.05 * torch.randn(b * v, 1, h, w)
    loss = sum((r["color"] * wc[i]).sum() + (r["depth"] * sc.near.reshape(-1)[i] * wd[i]).sum() for i, r in enumerate(ref))
    loss.backward()
    dec = _decoder(bg)
    g, t = _gaussians(sc, requires_grad=True)
    ext = sc.extrinsics.to(_dev()).requires_grad_()
    out = dec(g, ext, sc.intrinsics.to(_dev()), sc.near.to(_dev()), sc.far.to(_dev()), sc.image_shape)
    l2 = (out.color.reshape(b * v, 3, h, w) * wc.to(_dev())).sum() + (out.depth.reshape(b * v, 1, h, w) * wd.to(_dev())).sum()
    l2.backward()
    assert abs(l2.item() - loss.item()) <= 1e-4 * max(1.0, abs(loss.item()))
    for name in ("means", "scales", "rotations", "opacities", "harmonics"):
        assert rel_err(t[name].grad.cpu(), leaves[name].grad) < 1e-3, name
    assert rel_err(ext.grad.cpu(), leaves["extrinsics"].grad) < 1e-3


def test_backward_generations_agree():
    """Three independent implementations of the blend backward give the same gradients: the pair-log kernel (default:
    consumes the forward's log, closed form per pair), the recomputing pair-compaction kernel (pair_log=False, also the
    per-tile fallback when a warp's log overflows -- forced here with a tiny capacity) and the first-generation
    back-to-front kernel (SPF_FLAG_BWD_V1)."""
    from spfsplatv2_b200 import rasterizer as R
    from spfsplatv2_b200.camera import camera_setup
    from spfsplatv2_b200.rasterizer import RasterSettings, rasterize_batched
    d = _dev()
    for regime, h, w, grid in (("init", 128, 128, None), ("trained", 96, 80, (40, 40))):
        sc = make_scene(seed=29, v_cxt=1, h=h, w=w, grid=grid, regime=regime, n_target=2)
        view, proj, tanfov, scale = [x.to(d) for x in camera_setup(sc.extrinsics[0], sc.intrinsics[0], sc.near[0], sc.far[0], True)]
        wc = torch.randn(2, 3, h, w, device=d, generator=torch.Generator(device=d).manual_seed(1))
        grads, kinds = [], []
        for kind in ("log", "log_small_capacity", "recompute", "v1"):
            R._pair_cap_hint.clear(); R._pair_stat.clear()
            if kind == "log_small_capacity":     # most warps overflow -> their tiles fall back, the rest use the log
                R._pair_cap_hint[(0, 1, 2, sc.means.shape[1], h, w)] = 64
            t = {k: getattr(sc, k).to(d).requires_grad_() for k in ("means", "scales", "rotations", "opacities", "harmonics")}
            vm = view.clone().requires_grad_()
            s = RasterSettings(h, w, 4, 1.0, 2, sh_layout_ck=True, want_alpha=True, bwd_v1=(kind == "v1"),
                               pair_log=kind.startswith("log"))
            color, depth, alpha, _ = rasterize_batched(s, t["means"], t["scales"], t["rotations"], t["opacities"], t["harmonics"],
                                                       None, vm, proj, tanfov, torch.tensor([[0.3, 0.2, 0.1]] * 2, device=d), scale)
            ((color * wc).sum() + 0.1 * depth.sum() + 0.5 * (alpha * wc[:, :1]).sum()).backward()
            grads.append([t[k].grad for k in sorted(t)] + [vm.grad])
            kinds.append(kind)
        R._pair_cap_hint.clear(); R._pair_stat.clear()
        for kind, g in zip(kinds[1:], grads[1:]):
            for a, b in zip(grads[0], g):
                assert rel_err(a, b) < 2e-5, (regime, kind)


def test_pair_log_counts_and_capacity_feedback():
    """The forward's pair log: per-warp counts are consistent with n_contrib-derived totals, overflow is flagged, and the
    needed capacity is fed back to the next forward of the same shape."""
    from spfsplatv2_b200 import rasterizer as R
    sc = make_scene(seed=37, v_cxt=1, h=64, w=64, grid=(32, 32), regime="trained", n_target=1)
    from spfsplatv2_b200.camera import camera_setup
    from spfsplatv2_b200.rasterizer import RasterSettings, forward_with_state
    d = _dev()
    view, proj, tanfov, scale = [x.to(d) for x in camera_setup(sc.extrinsics[0], sc.intrinsics[0], sc.near[0], sc.far[0], True)]
    key = (0, 1, 1, sc.means.shape[1], 64, 64)
    args = (sc.means.to(d), sc.scales.to(d), sc.rotations.to(d), sc.opacities.to(d), sc.harmonics.to(d), None, view, proj, tanfov,
            torch.zeros(1, 3, device=d), scale)
    s = RasterSettings(64, 64, 4, 1.0, 1, sh_layout_ck=True)
    R._pair_cap_hint.clear(); R._pair_stat.clear()
    R._pair_cap_hint[key] = 4096
    st = forward_with_state(s, *args, pair_log=True)[4]
    counts = st.tensors["pair_count"].cpu()
    assert (counts >= 0).all()
    need = int(st.tensors["control"][2])
    assert need == int(counts.max()) and need > 64
    # total pairs == number of (pixel, Gaussian) contributions; cross-check against the oracle's blend weights
    ref, _ = oracle_views(sc)
    # every logged pair carries its pixel's lane and its record's index in the tile list; pairs of one record are
    # adjacent and in lane order; a pair is live iff its record index is below the pixel's n_contrib
    log = st.tensors["pair_log"].cpu().view(torch.int32).view(-1, 4096, 2)
    nc = st.tensors["n_contrib"][0].cpu()
    assert torch.equal(nc, ref[0]["n_contrib"])
    lens = (st.tensors["tile_ranges"][:, 1] - st.tensors["tile_ranges"][:, 0]).cpu()
    total = live = 0
    per_pixel = torch.zeros(64, 64, dtype=torch.int64)
    for wi in range(counts.numel()):
        c = int(counts[wi])
        if c == 0:
            continue
        w0 = log[wi, :c, 0].to(torch.int64) & 0xFFFFFFFF
        j = w0 & ((1 << 25) - 1)
        lane = (w0 >> 25) & 31
        tile, wid = wi // 8, wi % 8
        px = (tile % 4) * 16 + (wid % 2) * 8 + (lane % 8)
        py = (tile // 4) * 16 + (wid // 2) * 4 + (lane // 8)
        assert (j < int(lens[tile])).all()
        assert (j[1:] >= j[:-1]).all()                                   # list order
        same = j[1:] == j[:-1]
        assert (lane[1:][same] > lane[:-1][same]).all()                  # lane order inside a run
        alive = j < nc[py, px].to(torch.int64)
        per_pixel.index_put_((py[alive], px[alive]), torch.ones(int(alive.sum()), dtype=torch.int64), accumulate=True)
        total += c
        live += int(alive.sum())
    assert total > 0 and live > 0.9 * total
    # a pixel with n_contrib > 0 has at least one live pair and vice versa
    assert torch.equal(per_pixel > 0, nc > 0)
    R._pair_cap_hint[key] = 64
    st2 = forward_with_state(s, *args, pair_log=True)[4]
    c2 = st2.tensors["pair_count"].cpu()
    assert (c2 == -1).any() and int(st2.tensors["control"][2]) == need
    assert torch.equal(c2[c2 >= 0], counts[c2 >= 0])
    R._pair_cap_hint.clear(); R._pair_stat.clear()


def test_backward_is_bit_reproducible():
    sc = make_scene(seed=13, v_cxt=1, h=96, w=96, grid=(48, 48), regime="trained", n_target=2)
    dec = _decoder()
    grads = []
    for it in range(5):      # early iterations may grow the pair-log capacity (a different, equally valid summation order)
        g, t = _gaussians(sc, requires_grad=True)
        ext = sc.extrinsics.to(_dev()).requires_grad_()
        out = dec(g, ext, sc.intrinsics.to(_dev()), sc.near.to(_dev()), sc.far.to(_dev()), sc.image_shape)
        (out.color.square().sum() + out.depth.sum()).backward()
        grads.append([t[k].grad.clone() for k in sorted(t)] + [ext.grad.clone()])
    for a, b in zip(grads[3], grads[4]):
        assert torch.equal(a, b)
    for a, b in zip(grads[0], grads[4]):
        assert rel_err(a, b) < 1e-5


def test_colors_precomp_and_all_culled():
    from spfsplatv2_b200.decoder import render_cuda
    d = _dev()
    sc = make_scene(seed=17, v_cxt=1, h=64, w=64, grid=(16, 16), regime="trained", n_target=1, d_sh=1)
    # colours precomputed (use_sh=False): d_sh == 1 and the coefficient IS the colour
    ref, leaves = oracle_views(sc, requires_grad=True, use_sh=False)
    ref[0]["color"].square().sum().backward()
    args = [sc.extrinsics[0].to(d), sc.intrinsics[0].to(d), sc.near[0].to(d), sc.far[0].to(d), sc.image_shape,
            torch.zeros(1, 3, device=d)]
    harm = sc.harmonics.to(d).requires_grad_()
    means = sc.means.to(d).requires_grad_()
    img, dep = render_cuda(*args, means, sc.covariances.to(d), harm, sc.opacities.to(d), sc.rotations.to(d),
                           sc.scales.to(d), use_sh=False, enable_cov_grad=True, enable_sh_grad=True)
    assert (img[0].cpu() - ref[0]["color"]).abs().max().item() < 2e-5
    img.square().sum().backward()
    assert rel_err(harm.grad.cpu(), leaves["harmonics"].grad) < GRAD_TOL
    assert rel_err(means.grad.cpu(), leaves["means"].grad) < GRAD_TOL
    # everything behind the camera: empty lists, background only, zero gradients
    means_b = (sc.means * torch.tensor([1.0, 1.0, -1.0])).to(d).requires_grad_()
    bgc = torch.tensor([[0.25, 0.5, 0.75]], device=d)
    args[5] = bgc
    img, dep = render_cuda(*args, means_b, sc.covariances.to(d), sc.harmonics.to(d), sc.opacities.to(d),
                           sc.rotations.to(d), sc.scales.to(d), use_sh=False, enable_cov_grad=True,
                           enable_sh_grad=True)
    assert torch.allclose(img[0], bgc.view(3, 1, 1).expand(3, 64, 64)) and float(dep.abs().max()) == 0.0
    img.sum().backward()
    assert float(means_b.grad.abs().max()) == 0.0


def test_capacity_overflow_reruns():
    from spfsplatv2_b200 import rasterizer as R
    sc = make_scene(seed=19, v_cxt=1, h=64, w=64, grid=(32, 32), regime="trained", n_target=1)
    a = _state_forward(sc)
    R._capacity_hint.clear()
    key = (0, 1, 1, sc.means.shape[1], 64, 64)
    R._capacity_hint[key] = 1024          # far too small -> overflow flag -> re-run with the exact size
    assert a[4].n_dups > 1024
    b = _state_forward(sc)
    assert torch.equal(a[0], b[0]) and b[4].n_dups == a[4].n_dups


def _train_call(sc, view, proj, tanfov, scale, d):
    from spfsplatv2_b200.rasterizer import RasterSettings, rasterize_batched
    h, w = sc.image_shape
    t = {k: getattr(sc, k).to(d).requires_grad_() for k in ("means", "scales", "rotations", "opacities", "harmonics")}
    s = RasterSettings(h, w, 4, 1.0, 1, sh_layout_ck=True)
    color, depth, _, _ = rasterize_batched(s, t["means"], t["scales"], t["rotations"], t["opacities"], t["harmonics"], None,
                                           view, proj, tanfov, torch.zeros(1, 3, device=d), scale)
    return color, t


def test_training_overflow_is_never_silent(monkeypatch):
    """A training-mode forward sizes its duplicate buffers from earlier calls and does not wait for the count.  When the
    count jumps past the capacity (here: the hint is cut to half of what the scene needs) the image must be either
    correct (count already known: re-run inside the forward) or NaN-poisoned with the backward raising -- never a
    plausible picture with Gaussians missing; the step after the error is correct again."""
    from spfsplatv2_b200 import rasterizer as R
    from spfsplatv2_b200.camera import camera_setup
    d = _dev()
    sc = make_scene(seed=61, v_cxt=1, h=64, w=64, grid=(40, 40), regime="trained", n_target=1)
    view, proj, tanfov, scale = [x.to(d) for x in camera_setup(sc.extrinsics[0], sc.intrinsics[0], sc.near[0], sc.far[0], True)]
    key = (0, 1, 1, sc.means.shape[1], 64, 64)
    R._capacity_hint.clear(); R._unverified.clear()
    ref, t = _train_call(sc, view, proj, tanfov, scale, d)          # first call of the shape: exact
    n = R.last_forward_state().n_dups                               # (the state lives as long as the autograd graph)
    ref = ref.detach().clone()
    # (a) the count is already there when the forward looks: re-run inside the forward, image correct
    R._capacity_hint[key] = n // 2
    torch.cuda.synchronize()
    color, t = _train_call(sc, view, proj, tanfov, scale, d)
    torch.cuda.synchronize()
    if torch.isnan(color).any():        # the poll came too early on this run: must then behave like (b)
        with pytest.raises(R.DuplicateCapacityError):
            color.sum().backward()
    else:
        assert torch.equal(color.detach(), ref)
        color.sum().backward()
    # (b) the count is NOT there yet (forced): poisoned image, the backward raises, the retry is correct
    R._capacity_hint[key] = n // 2
    monkeypatch.setattr(R, "_poll_count", lambda *a: None)
    color, t = _train_call(sc, view, proj, tanfov, scale, d)
    assert bool(torch.isnan(color).all())
    with pytest.raises(R.DuplicateCapacityError):
        color.sum().backward()
    color, t = _train_call(sc, view, proj, tanfov, scale, d)
    assert torch.equal(color.detach(), ref)
    color.sum().backward()
    assert all(torch.isfinite(x.grad).all() for x in t.values())
    # (c) no backward ever follows the overflowed forward: the next forward of the shape reports it, the one after works
    R._capacity_hint[key] = n // 2
    color, t = _train_call(sc, view, proj, tanfov, scale, d)
    assert bool(torch.isnan(color).all())
    del color, t
    with pytest.raises(R.DuplicateCapacityError):
        _train_call(sc, view, proj, tanfov, scale, d)
    color, t = _train_call(sc, view, proj, tanfov, scale, d)
    assert torch.equal(color.detach(), ref)
    R._capacity_hint.clear(); R._unverified.clear()


def test_graph_replay_with_a_denser_scene_is_safe_and_flagged():
    """A captured step freezes the duplicate capacity.  Replaying it on a denser scene (scales x4 in place: far more
    tile duplicates than the capacity) must not write out of bounds (backward stores and reads are clipped to the
    capacity; run under compute-sanitizer in profiles/), yields a NaN loss and sets the overflow flag; restoring the
    scene and replaying reproduces the original loss bit for bit."""
    from spfsplatv2_b200 import rasterizer as R
    from spfsplatv2_b200.camera import camera_setup
    from spfsplatv2_b200.rasterizer import RasterSettings, rasterize_batched
    d = _dev()
    sc = make_scene(seed=67, v_cxt=1, h=64, w=64, grid=(40, 40), regime="trained", n_target=1)
    view, proj, tanfov, scale = [x.to(d) for x in camera_setup(sc.extrinsics[0], sc.intrinsics[0], sc.near[0], sc.far[0], True)]
    R._capacity_hint.clear(); R._pair_cap_hint.clear(); R._pair_stat.clear(); R._unverified.clear()
    stat = {k: getattr(sc, k).to(d) for k in ("means", "scales", "rotations", "opacities", "harmonics")}
    s = RasterSettings(64, 64, 4, 1.0, 1, sh_layout_ck=True)
    bg = torch.zeros(1, 3, device=d)

    def step():
        t = {k: v.detach().requires_grad_() for k, v in stat.items()}
        color, depth, _, _ = rasterize_batched(s, t["means"], t["scales"], t["rotations"], t["opacities"], t["harmonics"], None,
                                               view, proj, tanfov, bg, scale)
        loss = color.square().mean()
        loss.backward()
        return loss, t
    for _ in range(4):
        loss_e, _ = step()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss_g, t_g = step()
        st = R.last_forward_state()
    graph.replay(); torch.cuda.synchronize()
    assert torch.equal(loss_g, loss_e.detach()) and not R.graph_overflowed(st)
    small = stat["scales"].clone()
    stat["scales"].mul_(4.0)                       # in place: the graph reads the same memory
    graph.replay(); torch.cuda.synchronize()
    assert R.graph_overflowed(st) and bool(torch.isnan(loss_g))
    stat["scales"].copy_(small)
    graph.replay(); torch.cuda.synchronize()
    assert not R.graph_overflowed(st) and torch.equal(loss_g, loss_e.detach())
    assert all(torch.isfinite(x.grad).all() for x in t_g.values())


def test_shim_matches_batched_path():
    """The diff_gauss_pose drop-in (one view per call, [P,K,3] SH, python-float tanfov) gives the same
    image as the batched decoder path."""
    from spfsplatv2_b200.camera import camera_setup_cuda as camera_setup   # same kernel the decoder uses: same bits
    from spfsplatv2_b200.diff_gauss_pose import GaussianRasterizationSettings, GaussianRasterizer
    d = _dev()
    sc = make_scene(seed=23, v_cxt=1, h=64, w=48, grid=(32, 32), regime="trained", n_target=1)
    dec = _decoder()
    g, _ = _gaussians(sc)
    out = dec(g, sc.extrinsics.to(d), sc.intrinsics.to(d), sc.near.to(d), sc.far.to(d), sc.image_shape)
    view, proj, tanfov, scale = camera_setup(sc.extrinsics[0].to(d), sc.intrinsics[0].to(d), sc.near[0].to(d),
                                             sc.far[0].to(d), True)
    settings = GaussianRasterizationSettings(
        image_height=64, image_width=48, tanfovx=tanfov[0, 0].item(), tanfovy=tanfov[0, 1].item(),
        bg=torch.zeros(3, device=d), scale_modifier=1.0, projmatrix=proj[0], sh_degree=4, prefiltered=False,
        debug=False, enable_cov_grad=True, enable_sh_grad=True)
    means = (sc.means[0].to(d) * scale[0]).requires_grad_()
    m2d = torch.zeros_like(means, requires_grad=True)
    image, depth, norm, alpha, radii, extra = GaussianRasterizer(settings)(
        means3D=means, means2D=m2d, shs=sc.harmonics[0].permute(0, 2, 1).contiguous().to(d), colors_precomp=None,
        opacities=sc.opacities[0, :, None].to(d), scales=sc.scales[0].to(d) * scale[0],
        rotations=sc.rotations[0].to(d), viewmatrix=view[0])
    assert torch.equal(image, out.color[0, 0])
    assert radii.dtype == torch.int32 and radii.shape == (sc.means.shape[1],)
    image.sum().backward()
    assert m2d.grad is not None and means.grad is not None
    with pytest.raises(Exception):
        GaussianRasterizer(settings)(means3D=means, means2D=m2d, shs=None, colors_precomp=None,
                                     opacities=sc.opacities[0, :, None].to(d), scales=sc.scales[0].to(d),
                                     rotations=sc.rotations[0].to(d), viewmatrix=view[0])


def test_fused_camera_setup_matches_torch_glue():
    """csrc/camera.cu against the torch restatement of the reference's host glue, values and pose gradient."""
    from spfsplatv2_b200.camera import camera_setup, camera_setup_cuda
    d = _dev()
    torch.manual_seed(3)
    B = 5
    q = torch.randn(B, 4); q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                     2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1).view(B, 3, 3)
    ext = torch.eye(4).repeat(B, 1, 1); ext[:, :3, :3] = R; ext[:, :3, 3] = torch.randn(B, 3)
    K = torch.tensor([[0.88, 0, 0.5], [0, 0.91, 0.47], [0, 0, 1.0]]).repeat(B, 1, 1)
    near = torch.rand(B) + 0.1; far = near * 300
    for si in (True, False):
        e_cpu = ext.clone().requires_grad_()
        ref = camera_setup(e_cpu, K, near, far, si)
        e_gpu = ext.to(d).requires_grad_()
        got = camera_setup_cuda(e_gpu, K.to(d), near.to(d), far.to(d), si)
        for a, b in zip(got, ref):
            assert torch.allclose(a.cpu(), b, rtol=2e-5, atol=2e-5)
        wgt = torch.randn(B, 4, 4)
        (ref[0] * wgt).sum().backward()
        (got[0] * wgt.to(d)).sum().backward()
        assert rel_err(e_gpu.grad.cpu(), e_cpu.grad) < 1e-5


def test_cuda_graph_capture_of_a_whole_step():
    """fwd + fused loss + bwd captured in ONE CUDA graph after an eager step has sized the data-dependent buffers;
    replays reproduce the eager gradients bit for bit, and capturing a shape never seen eagerly is refused."""
    from spfsplatv2_b200 import rasterizer as R
    from spfsplatv2_b200.loss import mse_loss
    d = _dev()
    sc = make_batch(2, seed=47, v_cxt=1, h=64, w=64, grid=(40, 40), regime="trained", n_target=1)
    dec = _decoder()
    g, t = _gaussians(sc)
    ext = sc.extrinsics.to(d)
    gt = torch.rand(2, 1, 3, 64, 64, device=d)
    K, near, far = sc.intrinsics.to(d), sc.near.to(d), sc.far.to(d)
    from spfsplatv2_b200.decoder import Gaussians

    def step():
        leaves = {k: v.detach().requires_grad_() for k, v in t.items()}
        e = ext.detach().requires_grad_()
        gg = Gaussians(leaves["means"], g.covariances, leaves["rotations"], leaves["scales"], leaves["harmonics"], leaves["opacities"])
        out = dec(gg, e, K, near, far, sc.image_shape)
        loss = mse_loss(out.color, gt)
        loss.backward()
        return loss, leaves, e

    R._capacity_hint.clear(); R._pair_cap_hint.clear(); R._pair_stat.clear()
    graph = torch.cuda.CUDAGraph()
    with pytest.raises(RuntimeError, match="eager forward"):
        with torch.cuda.graph(graph):
            step()
    torch.cuda.synchronize()
    for _ in range(4):                       # eager steps: capacities (incl. the lagged pair-log feedback) settle
        loss_e, leaves_e, ext_e = step()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss_g, leaves_g, ext_g = step()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(loss_g, loss_e.detach())
    for k in leaves_e:
        assert torch.equal(leaves_g[k].grad, leaves_e[k].grad), k
    assert torch.equal(ext_g.grad, ext_e.grad)


@pytest.mark.parametrize("v_cxt,h,w", [(2, 256, 256), (2, 512, 512)])     # BASELINE configs 2 and 4 at full size
def test_full_size_properties(v_cxt, h, w):
    """Size-independent properties at BASELINE.json's full sizes (the oracle takes minutes there):
    keys sorted inside every tile and tile ranges partition the duplicate list; n_contrib within the tile list;
    alpha in [0,1], colour = blended + T*bg; the backward is linear in the upstream gradient and bit-reproducible;
    translation invariance ties the pose gradient to the mean gradients (sum_g dL/dm_g == dL/dtau A^T)."""
    from spfsplatv2_b200.camera import camera_setup
    from spfsplatv2_b200.rasterizer import RasterSettings, forward_with_state, rasterize_batched, unpack_sorted
    d = _dev()
    sc = make_scene(seed=53, v_cxt=v_cxt, h=h, w=w, regime="init", n_target=1)
    view, proj, tanfov, scale = [x.to(d) for x in camera_setup(sc.extrinsics[0], sc.intrinsics[0], sc.near[0], sc.far[0], True)]
    bg = torch.tensor([[0.2, 0.4, 0.6]], device=d)
    s = RasterSettings(h, w, 4, 1.0, 1, sh_layout_ck=True, want_alpha=True)
    base = (sc.means.to(d), sc.scales.to(d), sc.rotations.to(d), sc.opacities.to(d), sc.harmonics.to(d), None, view, proj, tanfov, bg, scale)
    color, depth, alpha, radii, st = forward_with_state(s, *base)
    N = st.n_dups
    assert N == int(st.tensors["tiles_touched"].sum())
    pl, keys = unpack_sorted(st)
    rg = st.tensors["tile_ranges"].long()
    T = rg.shape[0]
    lens = rg[:, 1] - rg[:, 0]
    assert int(lens.sum()) == N and int(lens.max()) > 0
    nz = lens > 0
    starts = rg[nz, 0]
    assert torch.equal(torch.sort(starts).values, starts)                       # ranges in tile order ...
    assert torch.equal(starts[1:], rg[nz, 1][:-1]) and int(starts[0]) == 0      # ... and contiguous: a partition
    tile_of = keys >> 32
    assert bool((tile_of[1:] >= tile_of[:-1]).all())                            # grouped by tile
    same = tile_of[1:] == tile_of[:-1]
    assert bool((keys[1:][same] >= keys[:-1][same]).all())                      # depth-sorted inside a tile
    eq = same & ((keys[1:] & 0xFFFFFFFF) == (keys[:-1] & 0xFFFFFFFF))
    assert bool((pl[1:][eq] > pl[:-1][eq]).all())                               # depth ties in Gaussian-id order
    gx = (w + 15) // 16
    tile_img = (torch.arange(h, device=d)[:, None] // 16) * gx + torch.arange(w, device=d)[None, :] // 16
    assert bool((st.tensors["n_contrib"][0].long() <= lens[tile_img]).all())
    assert float(alpha.min()) >= 0.0 and float(alpha.max()) <= 1.0
    acc = st.tensors["accum"][0]
    want = acc[..., :3].permute(2, 0, 1) + st.tensors["final_T"][0][None] * bg.view(3, 1, 1)
    assert torch.allclose(color[0], want, rtol=0, atol=2e-7)        # the kernel fuses T*bg + C into one FMA
    # backward: linearity, reproducibility, translation identity
    gen = torch.Generator(device=d).manual_seed(2)
    g1 = torch.randn(1, 3, h, w, device=d, generator=gen)
    g2 = torch.randn(1, 3, h, w, device=d, generator=gen)

    def grads(gc):
        t = [x.clone().requires_grad_() for x in base[:5]]
        vm = view.clone().requires_grad_()
        c, dd, a, _ = rasterize_batched(s, t[0], t[1], t[2], t[3], t[4], None, vm, proj, tanfov, bg, scale)
        c.backward(gc)
        return [x.grad for x in t] + [vm.grad]
    for _ in range(3):          # let the pair-log capacity settle (it changes which tiles use the log: different rounding)
        grads(g1)
    ga, gb, gab, ga2 = grads(g1), grads(g2), grads(g1 + 2.0 * g2), grads(g1)
    for x, y in zip(ga, ga2):
        assert torch.equal(x, y)
    for x, y, z in zip(ga, gb, gab):
        assert rel_err(z, x + 2.0 * y) < 2e-5
    # means enter the rasterizer as m*scale: dL/dm = scale * dL/d(m*scale)  ->  undo the chain factor
    lhs = ga[0][0].double().sum(0) / float(scale[0])
    rhs = ga[5][0, 3, :3].double() @ view[0, :3, :3].double().t()
    assert torch.allclose(lhs, rhs, rtol=2e-3, atol=2e-3 * float(lhs.abs().max()))


def test_odd_sizes_and_many_views_take_the_fallback_paths():
    """Shapes that cannot use the TMA-streamed projection kernels (P % 4 != 0 -> unaligned SH rows; more views than the
    per-CTA camera cache) run the one-shot kernels: same parity bars against the oracle, gradients summed over 36 views."""
    d = _dev()
    sc = make_scene(seed=59, v_cxt=1, h=48, w=40, grid=(27, 37), regime="trained", n_target=36)   # P = 999
    assert sc.means.shape[1] % 4 != 0
    ref, leaves = oracle_views(sc, bg=(0.1, 0.2, 0.3), requires_grad=True)
    wc = torch.randn(36, 3, 48, 40, generator=torch.Generator().manual_seed(4))
    loss = sum((r["color"] * wc[i]).sum() for i, r in enumerate(ref))
    loss.backward()
    color, depth, t, ext = _cuda_render_identical_inputs(sc, (0.1, 0.2, 0.3))
    for i in (0, 17, 35):
        assert (color[i].cpu() - ref[i]["color"]).abs().max().item() < 3e-5
    (color * wc.to(d)).sum().backward()
    for name in ("means", "scales", "rotations", "opacities", "harmonics"):
        e = rel_err(t[name].grad.cpu(), leaves[name].grad)
        assert e < GRAD_TOL, f"{name}: rel err {e:.3e}"
    assert rel_err(ext.grad, leaves["extrinsics"].grad) < GRAD_TOL


def test_many_views_on_the_streaming_projection_kernels():
    """More views than the per-CTA camera table holds (validation / video renders: up to 300 views of a scene) with
    shapes the TMA-streamed projection kernels accept: the view constants are then built per item.  Parity bars against
    the oracle on a few views, gradients summed over all 40."""
    d = _dev()
    sc = make_scene(seed=61, v_cxt=1, h=48, w=48, grid=(32, 32), regime="trained", n_target=40)    # P = 1024
    assert sc.means.shape[1] % 4 == 0
    ref, leaves = oracle_views(sc, bg=(0.2, 0.1, 0.3), requires_grad=True)
    wc = torch.randn(40, 3, 48, 48, generator=torch.Generator().manual_seed(6))
    loss = sum((r["color"] * wc[i]).sum() for i, r in enumerate(ref))
    loss.backward()
    color, depth, t, ext = _cuda_render_identical_inputs(sc, (0.2, 0.1, 0.3))
    for i in (0, 21, 33, 39):
        assert (color[i].cpu() - ref[i]["color"]).abs().max().item() < 3e-5
    (color * wc.to(d)).sum().backward()
    for name in ("means", "scales", "rotations", "opacities", "harmonics"):
        e = rel_err(t[name].grad.cpu(), leaves[name].grad)
        assert e < GRAD_TOL, f"{name}: rel err {e:.3e}"
    assert rel_err(ext.grad, leaves["extrinsics"].grad) < GRAD_TOL


def test_pose_align_graph_matches_the_eager_loop():
    """spfsplatv2_b200.pose_align.pose_align (the reference's test-time pose refinement, model_wrapper.py:539-590, as one
    captured CUDA graph per iteration) reproduces the plain eager loop -- decoder forward, MSE, backward, Adam on the
    extrinsics -- and moves a perturbed camera back towards the one the target image was rendered from."""
    from spfsplatv2_b200.decoder import Gaussians
    from spfsplatv2_b200.pose_align import pose_align
    d = _dev()
    sc = make_scene(seed=83, v_cxt=1, h=64, w=64, grid=(48, 48), regime="trained", n_target=2).to(d)
    dec = _decoder()
    g = Gaussians(sc.means, sc.covariances, sc.rotations, sc.scales, sc.harmonics, sc.opacities)
    with torch.no_grad():
        target = dec(g, sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape).color
    start = sc.extrinsics.clone()
    start[..., :3, 3] += torch.tensor([0.02, -0.015, 0.01], device=d)
    steps, lr = 24, 2e-3
    res = {}
    for use_graph in (False, True):
        out, ext, losses = pose_align(dec, g, start, sc.intrinsics, sc.near, sc.far, sc.image_shape, target, steps, lr,
                                      use_graph=use_graph, check_every=7)
        res[use_graph] = (out.color.clone(), ext.clone(), [float(x) for x in losses])
    assert res[True][2][-1] < 0.7 * res[True][2][0]                      # the loss went down
    err0 = (start - sc.extrinsics)[..., :3, 3].norm()
    err1 = (res[True][1] - sc.extrinsics)[..., :3, 3].norm()
    assert err1 < err0                                                   # and the camera moved towards the truth
    assert torch.allclose(res[True][1], res[False][1], rtol=0, atol=2e-5)           # same trajectory as the eager loop
    assert (res[True][0] - res[False][0]).abs().max().item() < 2e-3
    assert res[True][2][-1] == pytest.approx(res[False][2][-1], rel=1e-3)
