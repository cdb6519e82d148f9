"""The reference's OWN decoder code, unmodified, on top of this repo's drop-in rasterizer, on a GPU (SURVEY.md 8b:
integration level 0).  Needs a verbatim copy of the reference's `src` package at baseline/_ref/src (git-ignored; where
the base contract puts an installed reference) -- SPFSplatV2 is not pip-installable, so that is `cp -r
/root/reference/src baseline/_ref/src`.  The tree is NOT shipped with the repo (reference sources are never copied into
it): without it the tests skip and say so.  It was run once this round with the tree placed there by hand for that one
GPU call (profiles/r2_ref_decoder_on_dropin.log: both tests passed; sha256 of the two decoder files equal to
/root/reference's).  The only thing changed is what ``import diff_gauss_pose`` resolves to
(spfsplatv2_b200.install_shims, as INTEGRATION.md tells a maintainer to do).  Checked against (a) the fixture the same
reference code produced on the CPU oracle (tests/golden/decoder_ref.npz) and (b) this repo's batched
DecoderSplattingCUDA on the same inputs."""
import os
import sys

import numpy as np
import pytest
import torch

from tests.ref_probe import reference_tree, stub_absent_third_party_modules

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _reference_decoder_modules():
    tree = reference_tree()
    if tree is None:
        pytest.skip("baseline/_ref/src is absent (a verbatim copy of the reference's src package; never shipped with this "
                    "repo -- see the module docstring and profiles/r2_ref_decoder_on_dropin.log)")
    import spfsplatv2_b200
    stub_absent_third_party_modules()
    spfsplatv2_b200.install_shims()
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    sys.path.insert(0, tree)
    try:
        from src.model.decoder import get_decoder
        from src.model.decoder import cuda_splatting as ref_cs
        from src.model.decoder.decoder_splatting_cuda import DecoderSplattingCUDACfg
        from src.model.types import Gaussians
    finally:
        sys.path.remove(tree)
    assert "spfsplatv2_b200" in sys.modules["diff_gauss_pose"].__file__
    assert ref_cs.GaussianRasterizer.__module__.startswith("spfsplatv2_b200")       # the reference bound OUR rasterizer
    assert os.path.realpath(ref_cs.__file__).startswith(os.path.realpath(tree))
    return get_decoder, DecoderSplattingCUDACfg, Gaussians, ref_cs


def test_unmodified_reference_decoder_runs_on_the_drop_in_and_matches_its_own_cpu_fixture():
    get_decoder, Cfg, Gaussians, _ = _reference_decoder_modules()
    g = np.load(os.path.join(GOLD, "decoder_ref.npz"))
    dev = "cuda:0"
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    leaves = {k: t(k).requires_grad_() for k in ("means", "rotations", "scales", "harmonics", "opacities", "extrinsics")}
    h, w = [int(x) for x in g["image_shape"]]
    dec = get_decoder(Cfg("splatting_cuda", [float(x) for x in g["bg"]], True, True, True)).to(dev)
    b, P = g["means"].shape[:2]
    cov = torch.zeros(b, P, 3, 3, device=dev)
    out = dec.forward(Gaussians(leaves["means"], cov, leaves["rotations"], leaves["scales"], leaves["harmonics"],
                                leaves["opacities"]), leaves["extrinsics"], t("intrinsics"), t("near"), t("far"), (h, w))
    assert out.color.shape == g["color"].shape and out.depth.shape == g["depth"].shape
    ((out.color * t("wc")).sum() + (out.depth * t("wd")).sum()).backward()
    # same bars as the batched CUDA decoder gets against this fixture (tests/test_golden_gpu.py): the reference's glue runs
    # in torch on the GPU here, whose matrix inverse differs from the CPU's by ulps
    assert (out.color.detach().cpu() - torch.from_numpy(g["color"])).abs().max().item() < 1e-3
    assert (out.depth.detach().cpu() - torch.from_numpy(g["depth"])).abs().max().item() < 2e-3
    rel = lambda a, r: ((a.double() - r.double()).norm() / (r.double().norm() + 1e-30)).item()
    for k, leaf in leaves.items():
        assert rel(leaf.grad.cpu(), torch.from_numpy(g["grad_" + k])) < 1e-3, k


def test_unmodified_reference_render_cuda_agrees_with_the_batched_decoder():
    _, _, _, ref_cs = _reference_decoder_modules()
    from spfsplatv2_b200.decoder import render_cuda
    from spfsplatv2_b200.synthetic import make_batch
    dev = "cuda:0"
    sc = make_batch(3, seed=71, v_cxt=1, h=80, w=64, grid=(40, 40), regime="trained", n_target=1, with_cov=True).to(dev)
    B = 3
    bg = torch.tensor([0.1, 0.3, 0.2], device=dev).expand(B, 3)
    args = lambda L: (L["extrinsics"].reshape(B, 4, 4), sc.intrinsics.reshape(B, 3, 3), sc.near.reshape(B), sc.far.reshape(B),
                      (80, 64), bg, L["means"], sc.covariances, L["harmonics"], L["opacities"], L["rotations"], L["scales"])
    mk = lambda: {k: getattr(sc, k).clone().requires_grad_() for k in ("means", "rotations", "scales", "harmonics", "opacities", "extrinsics")}
    wc = torch.randn(B, 3, 80, 64, device=dev)
    L0, L1 = mk(), mk()
    c0, d0 = ref_cs.render_cuda(*args(L0), scale_invariant=True, use_sh=True, enable_cov_grad=True, enable_sh_grad=True)
    c1, d1 = render_cuda(*args(L1), scale_invariant=True, use_sh=True, enable_cov_grad=True, enable_sh_grad=True)
    (c0 * wc).sum().backward()
    (c1 * wc).sum().backward()
    assert c0.shape == c1.shape and d0.shape == d1.shape
    assert (c0 - c1).abs().max().item() < 1e-3 and (d0 - d1).abs().max().item() < 2e-3
    rel = lambda a, r: ((a.double() - r.double()).norm() / (r.double().norm() + 1e-30)).item()
    for k in L0:
        assert rel(L0[k].grad, L1[k].grad) < 1e-3, k
