"""Raw-head input of the projection kernels (SpfRasterIn.raw_head) against the unfused product path
-- stand-alone head kernel (pinned to the reference's adapter / opacity mapping by tests/golden/adapter_ref.npz in
test_golden_gpu.py) followed by the decoder (pinned to the oracle and the reference decoder fixture there) -- on the
same inputs.  Both paths share the per-element device functions (csrc/spf_adapter_math.cuh), so the images must be
identical bit for bit and the head gradients agree to rounding of the differently ordered view sums."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
D0 = "cuda:0"


def _inputs(seed, b, v, gh, gw, h, w, d_sh=25):
    from spfsplatv2_b200.synthetic import make_batch
    sc = make_batch(b, seed=seed, v_cxt=1, h=h, w=w, grid=(gh, gw), regime="init", n_target=v).to(D0)
    g = torch.Generator().manual_seed(seed + 77)
    P = sc.means.shape[1]
    head = torch.randn(b, P, 1 + 7 + 3 * d_sh, generator=g)
    head[..., 1:4] = head[..., 1:4] * 2.0 + 4.0          # scale logits: footprints of a few pixels, some at the 0.3 clamp
    head[:, :7, 1:4] = 400.0
    head[:, 7:9, 1:4] = 25.0                             # softplus threshold branch
    head[:, 9, 4:8] = 0.0                                # zero quaternion (eps path)
    return sc, head.to(D0)


def _decoder(cov=True, sh=True):
    from spfsplatv2_b200.decoder import DecoderSplattingCUDA, DecoderSplattingCUDACfg
    return DecoderSplattingCUDA(DecoderSplattingCUDACfg("splatting_cuda", [0.1, 0.2, 0.3], True, cov, sh)).to(D0)


def _loss(o, wc, wd):
    return (o.color * wc).sum() + (o.depth * wd).sum()


@pytest.mark.parametrize("b,v,exponent,flags", [(2, 1, 1.0, (True, True)), (1, 3, 2 ** 1.5, (True, True)),
                                                (2, 2, 1.0, (False, False)), (1, 36, 1.0, (True, True))])   # 36 > the camera table
def test_raw_head_matches_adapter_then_decoder(b, v, exponent, flags):
    from spfsplatv2_b200.adapter import GaussianAdapterCfg, OpacityMappingCfg, UnifiedGaussianAdapter
    sc, head = _inputs(3, b, v, 48, 48, 96, 96)
    dec = _decoder(*flags)
    ad = UnifiedGaussianAdapter(GaussianAdapterCfg(0.5, 15.0, 4))
    x = math.log2(exponent)
    cfg = OpacityMappingCfg(initial=x, final=x, warm_up=1)
    wc = torch.randn(b, v, 3, 96, 96, device=D0)
    wd = torch.randn(b, v, 96, 96, device=D0) * 0.1

    h0 = head.clone().requires_grad_()
    m0 = sc.means.clone().requires_grad_()
    e0 = sc.extrinsics.clone().requires_grad_()
    o0 = dec(ad.forward_head(m0, h0, cfg, 0), e0, sc.intrinsics, sc.near, sc.far, sc.image_shape)
    _loss(o0, wc, wd).backward()

    h1 = head.clone().requires_grad_()
    m1 = sc.means.clone().requires_grad_()
    e1 = sc.extrinsics.clone().requires_grad_()
    o1 = dec.forward_head(m1, h1, e1, sc.intrinsics, sc.near, sc.far, sc.image_shape, sh_degree=4, opacity_exponent=exponent)
    _loss(o1, wc, wd).backward()

    assert o0.color.abs().max().item() > 0.3                       # something was rendered
    assert torch.equal(o1.color, o0.color) and torch.equal(o1.depth, o0.depth)
    rel = lambda a, r: ((a.double() - r.double()).norm() / (r.double().norm() + 1e-30)).item()
    assert rel(m1.grad, m0.grad) < 1e-6 and rel(e1.grad, e0.grad) < 1e-6
    assert rel(h1.grad[..., 0], h0.grad[..., 0]) < 1e-6            # density logit through the opacity mapping
    assert rel(h1.grad[..., 1:4], h0.grad[..., 1:4]) < 1e-6        # scale logits
    assert rel(h1.grad[..., 4:8], h0.grad[..., 4:8]) < 1e-6        # raw quaternion
    assert rel(h1.grad[..., 8:], h0.grad[..., 8:]) < 1e-6          # SH coefficients x mask
    assert torch.isfinite(h1.grad).all()


def test_raw_head_with_separate_opacities_and_errors():
    """The adapter's own contract (rows of 7 + 3K, opacities given) and the entry point's argument checks."""
    from spfsplatv2_b200.adapter import GaussianAdapterCfg, UnifiedGaussianAdapter
    from spfsplatv2_b200.camera import camera_setup_cuda
    from spfsplatv2_b200.rasterizer import RasterSettings, rasterize_batched, rasterize_batched_head
    sc, head = _inputs(5, 2, 1, 32, 32, 64, 64)
    raw = head[..., 1:].contiguous()
    view, proj, tanfov, scale = camera_setup_cuda(sc.extrinsics.reshape(2, 4, 4), sc.intrinsics.reshape(2, 3, 3),
                                                  sc.near.reshape(2), sc.far.reshape(2), True)
    bg = torch.zeros(2, 3, device=D0)
    ad = UnifiedGaussianAdapter(GaussianAdapterCfg(0.5, 15.0, 4))
    r0 = raw.clone().requires_grad_()
    op0 = sc.opacities.clone().requires_grad_()
    gs = ad(sc.means, op0, r0)
    s_ck = RasterSettings(64, 64, 4, sh_layout_ck=True)
    c0, d0, _, rad0 = rasterize_batched(s_ck, gs.means, gs.scales, gs.rotations, gs.opacities, gs.harmonics, None, view, proj,
                                        tanfov, bg, scale)
    (c0.square().sum() + d0.sum()).backward()
    r1 = raw.clone().requires_grad_()
    op1 = sc.opacities.clone().requires_grad_()
    c1, d1, _, rad1 = rasterize_batched_head(RasterSettings(64, 64, 4), sc.means, r1, view, proj, tanfov, bg, scale, opacities=op1)
    (c1.square().sum() + d1.sum()).backward()
    assert torch.equal(c1, c0) and torch.equal(d1, d0) and torch.equal(rad1, rad0)
    rel = lambda a, r: ((a.double() - r.double()).norm() / (r.double().norm() + 1e-30)).item()
    assert rel(r1.grad, r0.grad) < 1e-6 and rel(op1.grad, op0.grad) < 1e-6

    with pytest.raises(RuntimeError, match="n_gaussians % 4"):       # P % 4 != 0
        rasterize_batched_head(RasterSettings(64, 64, 4), sc.means[:, :1023], head[:, :1023].contiguous(), view, proj, tanfov, bg)
    with pytest.raises(ValueError):       # row width does not match the SH degree
        rasterize_batched_head(RasterSettings(64, 64, 4), sc.means, head[..., :80].contiguous(), view, proj, tanfov, bg)


def test_half_precision_inputs_get_gradients_in_their_own_dtype():
    """bf16 head rows / Gaussian tensors (an autocast encoder) are widened to fp32 on the way in; the gradients come back
    in the input's dtype, as autograd requires."""
    from spfsplatv2_b200.camera import camera_setup_cuda
    from spfsplatv2_b200.rasterizer import RasterSettings, rasterize_batched, rasterize_batched_head
    sc, head = _inputs(9, 1, 1, 32, 32, 64, 64)
    view, proj, tanfov, scale = camera_setup_cuda(sc.extrinsics.reshape(1, 4, 4), sc.intrinsics.reshape(1, 3, 3),
                                                  sc.near.reshape(1), sc.far.reshape(1), True)
    bg = torch.zeros(1, 3, device=D0)
    h = head.to(torch.bfloat16).requires_grad_()
    c, d, _, _ = rasterize_batched_head(RasterSettings(64, 64, 4), sc.means, h, view, proj, tanfov, bg, scale)
    c.square().sum().backward()
    assert h.grad.dtype == torch.bfloat16 and torch.isfinite(h.grad.float()).all() and h.grad.float().abs().max().item() > 0
    m = sc.means.to(torch.float64).requires_grad_()
    sh = sc.harmonics.to(torch.float16).requires_grad_()
    c2, _, _, _ = rasterize_batched(RasterSettings(64, 64, 4, sh_layout_ck=True), m, sc.scales, sc.rotations, sc.opacities, sh, None,
                                    view, proj, tanfov, bg, scale)
    c2.sum().backward()
    assert m.grad.dtype == torch.float64 and sh.grad.dtype == torch.float16


def test_raw_head_with_a_lower_sh_degree():
    """Degree-2 harmonics (9 coefficients, rows of 1 + 7 + 27): the row stride, the SH mask and the bulk-copy sizes all
    depend on it."""
    from spfsplatv2_b200.adapter import GaussianAdapterCfg, UnifiedGaussianAdapter
    from spfsplatv2_b200.synthetic import make_batch
    sc = make_batch(2, seed=13, v_cxt=1, h=64, w=64, grid=(36, 36), regime="init", n_target=2, d_sh=9).to(D0)
    g = torch.Generator().manual_seed(5)
    head = torch.randn(2, sc.means.shape[1], 1 + 7 + 27, generator=g)
    head[..., 1:4] = head[..., 1:4] * 2.0 + 4.0
    head = head.to(D0)
    dec = _decoder()
    ad = UnifiedGaussianAdapter(GaussianAdapterCfg(0.5, 15.0, 2))
    h0, h1 = head.clone().requires_grad_(), head.clone().requires_grad_()
    o0 = dec(ad.forward_head(sc.means, h0), sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape)
    o1 = dec.forward_head(sc.means, h1, sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape, sh_degree=2)
    assert torch.equal(o0.color, o1.color) and torch.equal(o0.depth, o1.depth) and o0.color.abs().max().item() > 0.3
    wc = torch.randn_like(o0.color)
    (o0.color * wc).sum().backward()
    (o1.color * wc).sum().backward()
    rel = ((h1.grad.double() - h0.grad.double()).norm() / h0.grad.double().norm()).item()
    assert rel < 1e-6
