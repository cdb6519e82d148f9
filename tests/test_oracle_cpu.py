"""CPU tests: the oracle itself (BASELINE config 1), the shared per-Gaussian math (host build of
spf_math.h) against the oracle -- bit-exact on the index path, autograd-exact on the backward --
and the host glue against the reference's conventions."""
import ctypes
import math
import os

import pytest
import torch

from oracle import raster_oracle as O
from spfsplatv2_b200.camera import camera_setup, get_fov, get_projection_matrix
from spfsplatv2_b200.synthetic import make_scene
from tests.util import oracle_views, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cf = ctypes.c_float


def fp(t):
    return ctypes.c_void_p(t.data_ptr())


@pytest.fixture(scope="module")
def hostmath():
    so = os.path.join(ROOT, "tests", "hostmath", "libhostmath.so")
    if not os.path.exists(so):
        import subprocess
        subprocess.run(["bash", os.path.join(ROOT, "tests", "hostmath", "build.sh")], check=True)
    return ctypes.CDLL(so)


def _one_view(sc, h, w, requires_grad=False):
    view, proj, tanfov, scale = camera_setup(sc.extrinsics[0], sc.intrinsics[0], sc.near[0], sc.far[0], True)
    means = (sc.means[0] * scale[0]).contiguous()
    scales = (sc.scales[0] * scale[0]).contiguous()
    quats = sc.rotations[0].contiguous()
    shs = sc.harmonics[0].permute(0, 2, 1).contiguous()
    V, Pm = view[0].contiguous(), proj[0].contiguous()
    if requires_grad:
        for t in (means, scales, quats, shs, V):
            t.requires_grad_()
    vw = O.View(h, w, float(tanfov[0, 0]), float(tanfov[0, 1]), torch.zeros(3), V, Pm, 4, 1.0)
    return means, scales, quats, shs, V, Pm, tanfov, vw


def test_config1_oracle_renders_and_counts():
    """BASELINE config 1: 1k synthetic Gaussians -> 64x64, pure-PyTorch CPU alpha blend."""
    sc = make_scene(seed=0, v_cxt=1, h=64, w=64, grid=(32, 32), regime="init")
    res, _ = oracle_views(sc)
    r = res[0]
    assert r["color"].shape == (3, 64, 64) and r["depth"].shape == (1, 64, 64)
    assert torch.isfinite(r["color"]).all() and (r["alpha"] >= 0).all() and (r["alpha"] <= 1).all()
    n = int(r["pre"]["tiles_touched"].sum())
    assert n == r["keys"].numel() == r["point_list"].numel() and 900 < n < 2500
    # keys sorted, ranges partition the list, each tile's run carries that tile id
    assert torch.all(r["keys"][1:] >= r["keys"][:-1])
    for t in range(16):
        s, e = r["ranges"][t].tolist()
        assert torch.all((r["keys"][s:e] >> 32) == t)
    # front-to-back: depth non-decreasing inside a tile
    d = r["pre"]["depth"][r["point_list"].long()]
    s, e = r["ranges"][5].tolist()
    assert torch.all(d[s + 1:e] >= d[s:e - 1])


def test_oracle_blend_matches_scalar_loop():
    """The vectorised per-tile blend equals a literal per-pixel loop (SURVEY App. B 'Blend forward')."""
    sc = make_scene(seed=2, v_cxt=1, h=32, w=32, grid=(12, 12), regime="trained")
    res, _ = oracle_views(sc, bg=(0.1, 0.2, 0.3))
    r = res[0]
    pre = r["pre"]
    for (px, py) in [(3, 5), (17, 20), (31, 0), (16, 16)]:
        tile = (py // 16) * 2 + px // 16
        s, e = r["ranges"][tile].tolist()
        T, C, D, last, k = 1.0, torch.zeros(3), 0.0, 0, 0
        for i in range(s, e):
            g = int(r["point_list"][i]); k += 1
            dx = float(pre["xy"][g, 0]) - px; dy = float(pre["xy"][g, 1]) - py
            cx, cy, cz = [float(x) for x in pre["conic"][g]]
            power = -0.5 * (cx * dx * dx + cz * dy * dy) - cy * dx * dy
            if power > 0:
                continue
            a = min(0.99, float(pre["opacity"][g]) * math.exp(power))
            if a < 1 / 255:
                continue
            if T * (1 - a) < 1e-4:
                break
            C += pre["rgb"][g] * a * T; D += float(pre["depth"][g]) * a * T
            T *= (1 - a); last = k
        want = C + T * torch.tensor([0.1, 0.2, 0.3])
        assert torch.allclose(r["color"][:, py, px], want, atol=2e-5)
        assert abs(float(r["depth"][0, py, px]) - D) < 1e-3
        assert int(r["n_contrib"][py, px]) == last


def test_oracle_pose_gradient_finite_difference():
    """d/d(viewmatrix) through the (smooth) projection stage agrees with central differences; the blend
    itself has jump discontinuities (alpha cut-offs, tile membership), so FD is taken on the projected
    quantities with fixed random weights."""
    sc = make_scene(seed=4, v_cxt=1, h=32, w=32, grid=(10, 10), regime="trained")
    means, scales, quats, shs, V, Pm, tanfov, vw = _one_view(sc, 32, 32, requires_grad=True)
    torch.manual_seed(0)
    P = means.shape[0]
    wts = torch.randn(P, 9)

    def f(Vm):
        vw2 = O.View(32, 32, vw.tanfovx, vw.tanfovy, vw.bg, Vm, Pm, 4, 1.0)
        pre = O.preprocess(means.double().float(), scales, quats, sc.opacities[0], shs, None, vw2)
        vis = pre["visible"]
        q = torch.cat([pre["xy"], pre["conic"], pre["rgb"], pre["depth"][:, None]], dim=-1)
        return (q * wts)[vis].sum()
    (gV,) = torch.autograd.grad(f(V), V)
    eps = 1e-3
    for (i, j) in [(3, 0), (3, 1), (3, 2), (0, 0), (1, 2), (2, 1)]:
        dV = torch.zeros(4, 4); dV[i, j] = eps
        fd = (f(V.detach() + dV) - f(V.detach() - dV)).item() / (2 * eps)
        assert abs(fd - gV[i, j].item()) <= 0.02 * abs(gV[i, j].item()) + 0.05, (i, j, fd, gV[i, j].item())


def test_oracle_translation_invariance_identity():
    """Moving every Gaussian by D and the camera translation by -D*A leaves the image unchanged, hence
    sum_g dL/dm_g == dL/dtau * A^T exactly (up to fp32 summation).  Size-independent property that ties
    the pose gradient to the mean gradients; the GPU tests check the same identity at full size."""
    sc = make_scene(seed=8, v_cxt=1, h=48, w=48, grid=(16, 16), regime="trained")
    means, scales, quats, shs, V, Pm, tanfov, vw = _one_view(sc, 48, 48, requires_grad=True)
    torch.manual_seed(1)
    res = O.render(means, scales, quats, sc.opacities[0], shs, None, vw)
    (res["color"] * torch.randn(3, 48, 48)).sum().backward()
    lhs = means.grad.double().sum(0)
    rhs = V.grad[3, :3].double() @ V.detach()[:3, :3].double().t()
    assert torch.allclose(lhs, rhs, rtol=1e-3, atol=1e-3 * float(lhs.abs().max())), (lhs, rhs)


@pytest.mark.parametrize("regime", ["init", "trained"])
def test_shared_math_forward_bit_exact(hostmath, regime):
    sc = make_scene(seed=0, v_cxt=1, h=64, w=48, grid=(32, 32), regime=regime)
    means, scales, quats, shs, V, Pm, tanfov, vw = _one_view(sc, 64, 48)
    pre = O.preprocess(means, scales, quats, sc.opacities[0], shs, None, vw)
    P = means.shape[0]
    of = torch.zeros(P, 6); oi = torch.zeros(P, 6, dtype=torch.int32)
    hostmath.hm_project_forward(P, fp(means), fp(scales), fp(quats), fp(V), fp(Pm), cf(vw.tanfovx), cf(vw.tanfovy),
                                cf(1.0), 48, 64, fp(of), fp(oi))
    vis = pre["visible"]
    assert torch.equal(oi[:, 0], pre["radius"]) and torch.equal(oi[:, 1:5], pre["rect"])
    assert torch.equal(oi[:, 5], pre["tiles_touched"].to(torch.int32))
    bits = lambda t: t.contiguous().view(torch.int32)
    assert torch.equal(bits(of[:, 2]), bits(pre["depth"]))
    assert torch.equal(bits(of[vis, 0:2]), bits(pre["xy"][vis]))
    assert torch.equal(bits(of[vis, 3:6]), bits(pre["conic"][vis]))
    rgb = torch.zeros(P, 3)
    hostmath.hm_sh_forward(P, 4, fp(means), fp(V), fp(shs), fp(rgb))
    assert (rgb - pre["rgb"]).abs().max().item() < 1e-6


@pytest.mark.parametrize("regime", ["init", "trained"])
def test_shared_math_backward_matches_autograd(hostmath, regime):
    sc = make_scene(seed=1, v_cxt=1, h=64, w=48, grid=(32, 32), regime=regime)
    means, scales, quats, shs, V, Pm, tanfov, vw = _one_view(sc, 64, 48, requires_grad=True)
    pre = O.preprocess(means, scales, quats, sc.opacities[0], shs, None, vw)
    P = means.shape[0]
    torch.manual_seed(0)
    g2d = torch.randn(P, 10)
    vis = pre["visible"]
    g2d[~vis] = 0
    loss = ((pre["xy"] * g2d[:, 0:2]).sum() + (pre["conic"] * g2d[:, 2:5]).sum() + (pre["rgb"] * g2d[:, 6:9]).sum()
            + (pre["depth"] * g2d[:, 9]).sum())
    loss.backward()
    dm = torch.zeros(P, 3); ds = torch.zeros(P, 3); dq = torch.zeros(P, 4); dsh = torch.zeros(P, 25, 3); dV = torch.zeros(16)
    hostmath.hm_project_backward(P, 4, 1, 1, 1, fp(means.detach()), fp(scales.detach()), fp(quats.detach()),
                                 fp(shs.detach()), fp(V.detach()), fp(Pm), cf(vw.tanfovx), cf(vw.tanfovy), cf(1.0),
                                 48, 64, fp(g2d), fp(dm), fp(ds), fp(dq), fp(dsh), fp(dV))
    for mine, ref in ((dm, means.grad), (ds, scales.grad), (dq, quats.grad), (dsh, shs.grad)):
        mine[~vis] = 0
        r = ref.clone(); r[~vis] = 0
        assert rel_err(mine, r) < 1e-5
    assert rel_err(dV.view(4, 4), V.grad) < 1e-5
    assert float(V.grad[:, 3].abs().max()) == 0.0     # only V[:3,:3] and V[3,:3] carry gradient


def test_host_glue_conventions():
    """viewmatrix/projmatrix as the reference hands them over: transposed, translation in
    the last row, near -> 1 after the scale-invariance step."""
    sc = make_scene(seed=0, v_cxt=1, h=64, w=48, grid=(4, 4))
    ext = torch.eye(4)[None].clone()
    ext[0, :3, 3] = torch.tensor([-0.478, -0.281, 0.122])
    near, far = torch.tensor([0.1]), torch.tensor([100.0])
    view, proj, tanfov, scale = camera_setup(ext, sc.intrinsics[0], near, far, True)
    assert torch.allclose(view[0, 3, :3], torch.tensor([4.78, 2.81, -1.22]), atol=1e-5)
    assert torch.allclose(tanfov[0], torch.tensor([0.56818, 0.56818]), atol=1e-4)
    want = torch.tensor([[1.76, 0, 0, 0], [0, 1.76, 0, 0], [0, 0, 1.001, 1], [0, 0, -1.001, 0]])
    assert torch.allclose(proj[0], want, atol=2e-3)
    assert float(scale[0]) == pytest.approx(10.0)
    fov = get_fov(sc.intrinsics[0])
    assert math.degrees(float(fov[0, 0])) == pytest.approx(59.2, abs=0.1)
    p = get_projection_matrix(torch.tensor([1.0]), torch.tensor([1000.0]), fov[:, 0], fov[:, 1])
    assert p[0, 3, 2] == 1 and p[0, 2, 3] < 0
