"""CPU tests against the committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from
the reference's own code): the RoPE oracle against the reference's rope_2d_cpu and RoPE2D, and the host glue +
rasterizer oracle against the reference's unmodified DecoderSplattingCUDA.forward / render_cuda."""
import os

import numpy as np
import pytest
import torch

from oracle import rope_oracle as RO
from spfsplatv2_b200.camera import camera_setup
from spfsplatv2_b200.synthetic import Scene
from tests.util import oracle_views, rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROPE_CASES = ("small", "d8", "vit", "enc", "base10k")


@pytest.fixture(scope="module")
def rope_gold():
    return np.load(os.path.join(GOLD, "rope_ref.npz"))


@pytest.fixture(scope="module")
def dec_gold():
    return np.load(os.path.join(GOLD, "decoder_ref.npz"))


@pytest.mark.parametrize("case", ROPE_CASES)
def test_rope_oracle_matches_reference_cpp(rope_gold, case):
    """Vectorised oracle == the reference's rope_2d_cpu (curope.cpp:11-47) within 2 ulp-ish of fp32 powf/sincos
    differences between glibc (reference build) and numpy; forward and backward (-F0)."""
    tok, pos, base = rope_gold[f"{case}_tokens"], rope_gold[f"{case}_pos"], float(rope_gold[f"{case}_base"])
    for key, f0 in (("fwd", 1.0), ("bwd", -1.0)):
        got = RO.rope_2d(tok, pos, base, f0)
        assert np.abs(got - rope_gold[f"{case}_{key}"]).max() < 2e-6
    # the reference's pure-PyTorch fallback (pos_embed.py:112-159) is the same function
    assert np.abs(RO.rope_2d(tok, pos, base, 1.0) - rope_gold[f"{case}_pytorch_fwd"]).max() < 5e-6


def test_rope_loop_restatement_matches_vectorised(rope_gold):
    for case in ("small", "d8"):
        tok, pos, base = rope_gold[f"{case}_tokens"], rope_gold[f"{case}_pos"], float(rope_gold[f"{case}_base"])
        a = RO.rope_2d_loop(tok, pos, base, 1.0)
        assert np.abs(a - RO.rope_2d(tok, pos, base, 1.0)).max() < 1e-6
        assert np.abs(a - rope_gold[f"{case}_fwd"]).max() < 2e-6


def test_rope_properties(rope_gold):
    """Size-independent properties: round trip (+F0 then -F0) is the identity, rotations preserve the norm of
    every (u,v) pair, position 0 is the identity."""
    tok, pos, base = rope_gold["vit_tokens"], rope_gold["vit_pos"], float(rope_gold["vit_base"])
    f = RO.rope_2d(tok, pos, base, 1.0)
    assert np.abs(RO.rope_2d(f, pos, base, -1.0) - tok).max() < 2e-6
    Q = tok.shape[-1] // 4
    def pair_norm(t):
        t = t.reshape(*t.shape[:-1], 2, 2, Q)
        return (t ** 2).sum(-2)
    assert np.allclose(pair_norm(f), pair_norm(tok), rtol=1e-5, atol=1e-6)
    assert np.array_equal(RO.rope_2d(tok, np.zeros_like(pos), base, 1.0), tok)


def _scene(g) -> Scene:
    t = lambda k: torch.from_numpy(g[k])
    P = g["means"].shape[1]
    return Scene(t("means"), torch.zeros(g["means"].shape[0], P, 3, 3), t("rotations"), t("scales"), t("harmonics"),
                 t("opacities"), t("extrinsics"), t("intrinsics"), t("near"), t("far"), tuple(int(x) for x in g["image_shape"]))


def test_camera_glue_matches_reference_render_cuda(dec_gold):
    """camera_setup reproduces, bit for bit, every per-view argument the reference's render_cuda hands to the
    rasterizer (cuda_splatting.py:66-90,105-138): viewmatrix, projmatrix, tanfov, and the 1/near scaling."""
    g = dec_gold
    sc = _scene(g)
    b, v = sc.extrinsics.shape[:2]
    view, proj, tanfov, scale = camera_setup(sc.extrinsics.reshape(b * v, 4, 4), sc.intrinsics.reshape(b * v, 3, 3),
                                             sc.near.reshape(-1), sc.far.reshape(-1), True)
    assert np.array_equal(view.numpy(), g["rec_viewmatrix"])
    assert np.array_equal(proj.numpy(), g["rec_projmatrix"])
    assert np.array_equal(tanfov.double().numpy(), g["rec_tanfov"])     # .item() of fp32 -> python float
    assert not g["rec_proj_contiguous"].any()                            # the shim must accept strided projmatrix
    assert (g["rec_sh_degree"] == 4).all() and (g["rec_shs_shape"] == [576, 25, 3]).all()
    assert (g["rec_opac_shape"] == [576, 1]).all()
    for i in range(b * v):
        s = i // v
        assert (sc.means[s] * scale[i]).double().sum().item() == pytest.approx(float(g["rec_means_sum"][i]), rel=1e-12)
        assert (sc.scales[s] * scale[i]).double().sum().item() == pytest.approx(float(g["rec_scales_sum"][i]), rel=1e-12)
        assert np.array_equal(g["rec_bg"][i], g["bg"])


def test_oracle_path_matches_reference_decoder_forward_backward(dec_gold):
    """tests.util.oracle_views (our glue + oracle) == the reference's unmodified DecoderSplattingCUDA.forward driving
    the same oracle rasterizer: outputs and gradients (incl. camera pose), so the harness the GPU parity tests use
    is equivalent to the reference's own host code."""
    g = dec_gold
    sc = _scene(g)
    b, v = sc.extrinsics.shape[:2]
    res, leaves = oracle_views(sc, bg=tuple(float(x) for x in g["bg"]), requires_grad=True)
    color = torch.stack([r["color"] for r in res]).view(b, v, 3, *sc.image_shape)
    depth = torch.stack([r["depth"][0] for r in res]).view(b, v, *sc.image_shape) * sc.near[:, :, None, None]
    assert np.abs(color.detach().numpy() - g["color"]).max() < 1e-6
    assert np.abs(depth.detach().numpy() - g["depth"]).max() < 1e-5
    loss = (color * torch.from_numpy(g["wc"])).sum() + (depth * torch.from_numpy(g["wd"])).sum()
    assert loss.item() == pytest.approx(float(g["loss"]), rel=1e-5)
    loss.backward()
    for k in ("means", "rotations", "scales", "harmonics", "opacities", "extrinsics"):
        assert rel_err(leaves[k].grad, torch.from_numpy(g["grad_" + k])) < 2e-6, k


def test_orthographic_glue_matches_reference():
    """orthographic_setup == the camera math of the reference's render_cuda_orthographic (cuda_splatting.py:173-202),
    bit for bit, including the tensor-valued tanfov it hands to the rasterizer settings (:221-222)."""
    from spfsplatv2_b200.camera import orthographic_setup
    g = np.load(os.path.join(GOLD, "ortho_ref.npz"))
    t = lambda k: torch.from_numpy(g[k])
    view, proj, tanfov, info = orthographic_setup(t("extrinsics"), t("width"), t("height"), t("near"), t("far"), 0.1)
    assert np.array_equal(view[0].numpy(), g["rec_viewmatrix"])
    assert np.array_equal(proj[0].numpy(), g["rec_projmatrix"])
    assert np.array_equal(tanfov[0].double().numpy(), g["rec_tanfov"])
    assert list(g["tanfov_types"]) == ["Tensor", "Tensor"]          # the shim must coerce tensors to floats
    assert np.array_equal(info["near"].numpy(), g["dump_near"]) and np.array_equal(info["far"].numpy(), g["dump_far"])


def test_fullsize_fixture_c2p_reproduced_by_the_oracle():
    """tests/golden/fullsize_c2p.npz (the headline 65k scene at 256x256) is what the oracle produces today on the seeded
    scene: guards the fixture against drift of the oracle, of the scene generator and of torch's CPU RNG."""
    import os
    from tests.golden import fullsize as F
    from tests.util import oracle_views
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize_c2p.npz")
    fx = np.load(p)
    sc = F.scene_of("c2p")
    assert F.inputs_digest(sc) == str(fx["inputs_sha"])
    res, _ = oracle_views(sc, bg=F.CONFIGS["c2p"]["bg"])
    r = res[0]
    assert r["keys"].numel() == int(fx["n_dups"])
    assert F.sha(r["keys"]) == str(fx["keys_sha"]) and F.sha(r["point_list"]) == str(fx["point_list_sha"])
    assert torch.equal(r["n_contrib"], torch.from_numpy(fx["n_contrib"].astype(np.int32)))
    assert torch.equal(r["color"], torch.from_numpy(fx["color"]))
    for name in F.CONFIGS:      # every committed fixture belongs to the scene its config describes
        q = os.path.join(os.path.dirname(p), f"fullsize_{name}.npz")
        if os.path.exists(q):
            assert F.inputs_digest(F.scene_of(name)) == str(np.load(q)["inputs_sha"]), name
