#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/ by running the REFERENCE's own code in this
container (it cannot travel to the GPU box; the fixtures can).  Test infrastructure only.

    python tests/golden/make_golden.py            # needs /root/reference and oracle/_ref (oracle/build_ref.sh)

Fixtures
  rope_ref.npz      outputs of the reference's rope_2d_cpu (curope.cpp:11-47, compiled unmodified into
                    oracle/_ref/curope_ref*.so) and of its pure-PyTorch RoPE2D fallback (pos_embed.py:112-159)
                    on seeded tokens / positions, forward (+F0) and backward (-F0).
  adapter_ref.npz   the reference's UnifiedGaussianAdapter (gaussian_adapter.py:122-150): outputs and d/d(raw) of a seeded loss.
  ortho_ref.npz     the reference's render_cuda_orthographic (cuda_splatting.py:146-255) on a seeded box of Gaussians, same
                    recording stand-in for the rasterizer: image + the arguments it passes (tensor-valued tanfov).
  decoder_ref.npz   the reference's UNMODIFIED DecoderSplattingCUDA.forward (decoder_splatting_cuda.py:41-78) and
                    render_cuda (cuda_splatting.py:45-144) driven end to end on a seeded scene, with the external
                    diff_gauss_pose package (absent) replaced by a recording module whose rasterizer
                    is oracle/raster_oracle.py.  Holds the inputs, every per-view argument the reference hands to
                    the rasterizer (viewmatrix, projmatrix, tanfov, bg, scaled means / scales checksums), the
                    decoder outputs (color, depth) and the gradients of a seeded loss wrt all Gaussian inputs and
                    the camera extrinsics.  This pins the reference's host glue exactly; the rasterizer arithmetic
                    itself stays "parity unpinned" (no reference implementation or vectors exist for it).
"""
from __future__ import annotations

import math
import os
import sys
import types
from typing import NamedTuple

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


# ------------------------------------------------------------------------------------------------ RoPE
def make_rope():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import curope_ref  # the reference's curope.cpp, CPU path
    # the reference's pure-torch fallback; importing pos_embed tries `.curope` first and falls back on ImportError
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "ref_pos_embed", os.path.join(REF, "src/model/encoder/backbone/croco/pos_embed.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.RoPE2D.__name__ == "RoPE2D" and hasattr(mod.RoPE2D, "apply_rope1d"), "expected the torch fallback"

    # VGGT's own pure-torch RoPE (vggt/layers/rope.py:62-188), loaded from where it lies
    vspec = importlib.util.spec_from_file_location(
        "ref_vggt_rope", os.path.join(REF, "src/model/encoder/backbone/vggt/layers/rope.py"))
    vmod = importlib.util.module_from_spec(vspec)
    vspec.loader.exec_module(vmod)

    out = {}
    cases = {   # name: (B, N, H, D, max_pos, base)
        "small": (2, 7, 3, 16, 5, 100.0),
        "d8": (1, 5, 2, 8, 40, 100.0),            # Q=2: scalar path of the kernel
        "vit": (2, 37, 12, 64, 16, 100.0),        # decoder head shape (12 heads x 64), 16x16 patch grid
        "enc": (1, 20, 16, 64, 16, 100.0),        # encoder head shape
        "base10k": (1, 9, 2, 32, 30, 10000.0),
    }
    g = torch.Generator().manual_seed(0)
    for name, (B, N, H, D, mp, base) in cases.items():
        tok = torch.randn(B, N, H, D, generator=g)
        pos = torch.randint(0, mp, (B, N, 2), generator=g, dtype=torch.int64)
        fwd = tok.clone()
        curope_ref.rope_2d(fwd, pos, base, 1.0)
        bwd = tok.clone()
        curope_ref.rope_2d(bwd, pos, base, -1.0)
        rt = fwd.clone()
        curope_ref.rope_2d(rt, pos, base, -1.0)
        py = mod.RoPE2D(freq=base, F0=1.0)(tok.transpose(1, 2).clone(), pos).transpose(1, 2).contiguous()
        out[f"{name}_tokens"] = tok.numpy()
        out[f"{name}_pos"] = pos.numpy()
        out[f"{name}_base"] = np.float32(base)
        out[f"{name}_fwd"] = fwd.numpy()
        out[f"{name}_bwd"] = bwd.numpy()
        out[f"{name}_pytorch_fwd"] = py.numpy()
        if D % 2 == 0:
            # VGGT layout: tokens [B, H, N, D]; its positions come from PositionGetter (0-based grid coordinates)
            vg = vmod.RotaryPositionEmbedding2D(frequency=base)(tok.transpose(1, 2).contiguous(), pos)
            out[f"{name}_vggt_fwd"] = vg.transpose(1, 2).contiguous().numpy()
            print(f"rope {name}: vggt-vs-cpp max|d| = {(out[f'{name}_vggt_fwd'] - fwd.numpy()).__abs__().max():.2e}")
        print(f"rope {name}: cpp-vs-pytorch max|d| = {(fwd - py).abs().max():.2e}, round trip {(rt - tok).abs().max():.2e}")
    np.savez_compressed(os.path.join(HERE, "rope_ref.npz"), **out)


# --------------------------------------------------------------------------------------------- decoder
def _stub_modules():
    import torchvision  # noqa: F401  (real one first)

    class _Any(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            t = type(name, (), {"__init__": lambda self, *a, **k: None})
            setattr(self, name, t)
            return t

    for name in ("dacite", "lightning", "lightning.pytorch", "skvideo", "skvideo.io", "matplotlib",
                 "matplotlib.figure", "omegaconf", "pytorch3d", "pytorch3d.transforms", "lightning.pytorch.loggers",
                 "lightning.pytorch.loggers.wandb", "lightning.pytorch.utilities", "lightning.pytorch.callbacks",
                 "lightning.pytorch.plugins.environments", "lightning.pytorch.plugins", "matplotlib.pyplot",
                 "matplotlib.cm", "matplotlib.colors", "lpips", "plyfile", "wandb", "colorspacious", "moviepy",
                 "moviepy.editor", "hydra", "skimage", "skimage.metrics", "roma", "e3nn", "e3nn.o3", "timm", "timm.models",
                 "timm.models.layers", "timm.layers", "huggingface_hub", "safetensors", "safetensors.torch", "xformers", "xformers.ops"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = _Any(name)
                m.__path__ = []
                sys.modules[name] = m


RECORD = []


def _fake_diff_gauss_pose():
    from oracle import raster_oracle as O

    class GaussianRasterizationSettings(NamedTuple):
        image_height: int
        image_width: int
        tanfovx: float
        tanfovy: float
        bg: torch.Tensor
        scale_modifier: float
        projmatrix: torch.Tensor
        sh_degree: int
        prefiltered: bool
        debug: bool
        enable_cov_grad: bool
        enable_sh_grad: bool

    class GaussianRasterizer(torch.nn.Module):
        def __init__(self, raster_settings):
            super().__init__()
            self.s = raster_settings

        def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                    cov3Ds_precomp=None, viewmatrix=None):
            s = self.s
            RECORD.append(dict(
                tanfov=(s.tanfovx, s.tanfovy), tanfov_types=(type(s.tanfovx).__name__, type(s.tanfovy).__name__),
                bg=s.bg.detach().clone(), projmatrix=s.projmatrix.detach().clone(), viewmatrix=viewmatrix.detach().clone(),
                proj_contiguous=s.projmatrix.is_contiguous(), sh_degree=s.sh_degree, hw=(s.image_height, s.image_width),
                means_sum=means3D.detach().double().sum().item(), scales_sum=scales.detach().double().sum().item(),
                shs_shape=None if shs is None else tuple(shs.shape), opac_shape=tuple(opacities.shape),
                flags=(s.prefiltered, s.debug, s.enable_cov_grad, s.enable_sh_grad, s.scale_modifier)))
            vw = O.View(s.image_height, s.image_width, float(s.tanfovx), float(s.tanfovy), s.bg,
                        viewmatrix.contiguous(), s.projmatrix.contiguous(), s.sh_degree, s.scale_modifier)
            r = O.render(means3D, scales, rotations, opacities.reshape(-1), shs, colors_precomp, vw)
            return r["color"], r["depth"], None, r["alpha"], r["pre"]["radius"], None

    m = types.ModuleType("diff_gauss_pose")
    m.GaussianRasterizationSettings = GaussianRasterizationSettings
    m.GaussianRasterizer = GaussianRasterizer
    return m


def make_decoder():
    _stub_modules()
    sys.modules["diff_gauss_pose"] = _fake_diff_gauss_pose()
    sys.path.insert(0, REF)
    from src.model.decoder import get_decoder
    from src.model.decoder.decoder_splatting_cuda import DecoderSplattingCUDACfg
    from src.model.types import Gaussians
    from spfsplatv2_b200.synthetic import make_batch

    b, v, h, w = 2, 2, 64, 48
    bg = [0.2, 0.1, 0.4]
    sc = make_batch(b, seed=31, v_cxt=1, h=h, w=w, grid=(24, 24), regime="trained", n_target=v, with_cov=True)
    # vary near per view and use an off-centre principal point so the glue's handling of both is pinned
    sc.near = torch.tensor([[0.5, 0.8], [0.3, 1.0]])
    sc.far = sc.near * 1000.0
    sc.intrinsics = sc.intrinsics.clone()
    sc.intrinsics[1, :, 0, 2] = 0.47
    sc.intrinsics[1, :, 1, 1] = 0.91
    leaves = {k: getattr(sc, k).clone().requires_grad_() for k in
              ("means", "rotations", "scales", "harmonics", "opacities", "extrinsics")}
    dec = get_decoder(DecoderSplattingCUDACfg("splatting_cuda", bg, True, True, True))
    g = Gaussians(leaves["means"], sc.covariances, leaves["rotations"], leaves["scales"], leaves["harmonics"],
                  leaves["opacities"])
    out = dec.forward(g, leaves["extrinsics"], sc.intrinsics, sc.near, sc.far, (h, w))
    gen = torch.Generator().manual_seed(1)
    wc = torch.randn(b, v, 3, h, w, generator=gen)
    wd = 0.05 * torch.randn(b, v, h, w, generator=gen)
    loss = (out.color * wc).sum() + (out.depth * wd).sum()
    loss.backward()
    assert len(RECORD) == b * v
    assert all(r["tanfov_types"] == ("float", "float") for r in RECORD)
    fx = dict(
        bg=np.array(bg, np.float32), image_shape=np.array([h, w]),
        means=sc.means.numpy(), rotations=sc.rotations.numpy(), scales=sc.scales.numpy(),
        harmonics=sc.harmonics.numpy(), opacities=sc.opacities.numpy(), extrinsics=sc.extrinsics.numpy(),
        intrinsics=sc.intrinsics.numpy(), near=sc.near.numpy(), far=sc.far.numpy(),
        wc=wc.numpy(), wd=wd.numpy(), loss=np.float64(loss.item()),
        color=out.color.detach().numpy(), depth=out.depth.detach().numpy(),
        rec_viewmatrix=torch.stack([r["viewmatrix"] for r in RECORD]).numpy(),
        rec_projmatrix=torch.stack([r["projmatrix"] for r in RECORD]).numpy(),
        rec_tanfov=np.array([r["tanfov"] for r in RECORD], np.float64),
        rec_bg=torch.stack([r["bg"] for r in RECORD]).numpy(),
        rec_means_sum=np.array([r["means_sum"] for r in RECORD]),
        rec_scales_sum=np.array([r["scales_sum"] for r in RECORD]),
        rec_proj_contiguous=np.array([r["proj_contiguous"] for r in RECORD]),
        rec_sh_degree=np.array([r["sh_degree"] for r in RECORD]),
        rec_shs_shape=np.array([r["shs_shape"] for r in RECORD]),
        rec_opac_shape=np.array([r["opac_shape"] for r in RECORD]),
    )
    for k, t in leaves.items():
        fx["grad_" + k] = t.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "decoder_ref.npz"), **fx)
    print(f"decoder: {len(RECORD)} rasterizer calls, loss {loss.item():.6f}, color mean {out.color.mean().item():.4f}; "
          f"proj contiguous={RECORD[0]['proj_contiguous']}, shs {RECORD[0]['shs_shape']}, opac {RECORD[0]['opac_shape']}")


def make_orthographic():
    """The reference's render_cuda_orthographic (cuda_splatting.py:146-255; B = 1, the only batch size its
    `move_back[2, 3] = -distance_to_near` / [B]-shaped tanfovy arguments are sane for) on top of the oracle."""
    _stub_modules()
    RECORD.clear()
    sys.modules["diff_gauss_pose"] = _fake_diff_gauss_pose()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from src.model.decoder.cuda_splatting import render_cuda_orthographic
    g = torch.Generator().manual_seed(7)
    P, h, w = 400, 48, 40
    means = (torch.rand(1, P, 3, generator=g) - 0.5) * torch.tensor([2.4, 2.8, 2.0])
    scales = 0.02 + 0.06 * torch.rand(1, P, 3, generator=g)
    rot = torch.randn(1, P, 4, generator=g)
    rot = rot / rot.norm(dim=-1, keepdim=True)
    opac = torch.sigmoid(torch.randn(1, P, generator=g))
    harm = torch.randn(1, P, 3, 25, generator=g) * 0.3
    ext = torch.eye(4)[None].clone()
    ext[0, :3, 3] = torch.tensor([0.1, -0.05, -3.0])
    width, height = torch.tensor([3.0]), torch.tensor([3.6])
    near, far = torch.tensor([0.5]), torch.tensor([20.0])
    bg = torch.tensor([[0.1, 0.2, 0.3]])
    dump = {}
    img = render_cuda_orthographic(ext, width, height, near, far, (h, w), bg, means, torch.zeros(1, P, 3, 3), harm, opac,
                                   rot, scales, fov_degrees=0.1, use_sh=True, dump=dump)
    r = RECORD[0]
    np.savez_compressed(os.path.join(HERE, "ortho_ref.npz"), means=means.numpy(), scales=scales.numpy(), rotations=rot.numpy(),
                        opacities=opac.numpy(), harmonics=harm.numpy(), extrinsics=ext.numpy(), width=width.numpy(),
                        height=height.numpy(), near=near.numpy(), far=far.numpy(), bg=bg.numpy(), image_shape=np.array([h, w]),
                        image=img.detach().numpy(), rec_viewmatrix=r["viewmatrix"].numpy(), rec_projmatrix=r["projmatrix"].numpy(),
                        rec_tanfov=np.array([float(r["tanfov"][0]), float(r["tanfov"][1])]),
                        tanfov_types=np.array(r["tanfov_types"]), dump_near=dump["near"].numpy(), dump_far=dump["far"].numpy())
    print(f"orthographic: image mean {img.mean().item():.4f}, tanfov types {r['tanfov_types']}, covered {(img[0] != bg.view(3,1,1)).any(0).float().mean():.2f}")


def make_adapter():
    """The reference's UnifiedGaussianAdapter (gaussian_adapter.py:122-150), forward and backward, on seeded raw head
    outputs incl. scale logits beyond the softplus threshold (20) and the 0.3 clamp."""
    _stub_modules()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from src.model.encoder.common.gaussian_adapter import GaussianAdapterCfg, UnifiedGaussianAdapter
    ad = UnifiedGaussianAdapter(GaussianAdapterCfg(gaussian_scale_min=0.5, gaussian_scale_max=15.0, sh_degree=4))
    g = torch.Generator().manual_seed(11)
    b, n = 2, 301
    raw = torch.randn(b, n, 82, generator=g)
    raw[0, :8, 0] = torch.tensor([25.0, 19.9, 20.1, -30.0, 350.0, 299.9, 300.5, 0.0])     # softplus threshold / 0.3 clamp
    raw[1, 0, 3:7] = 0.0                                                               # zero quaternion (eps path)
    raw = raw.requires_grad_()
    means = torch.randn(b, n, 3, generator=g)
    opac = torch.rand(b, n, generator=g)
    out = ad.forward(means, opac, raw)
    ws, wr, wh = torch.randn(b, n, 3, generator=g), torch.randn(b, n, 4, generator=g), torch.randn(b, n, 3, 25, generator=g)
    ((out.scales * ws).sum() + (out.rotations * wr).sum() + (out.harmonics * wh).sum()).backward()
    extra = {}
    # The encoder's opacity mapping (EncoderSPFSplatV2.map_pdf_to_opacity, encoder_spfsplatv2.py:146-159) run from the
    # reference file itself: the method is lifted out of the class by its AST (importing the module would pull in the whole
    # backbone) and called on a stand-in `self` that only carries cfg.opacity_mapping.  Then the head post-processing
    # of :255-268: densities = sigmoid(channel 0) -> opacities; the rest -> the adapter.
    import ast
    import textwrap
    src = open(os.path.join(REF, "src/model/encoder/encoder_spfsplatv2.py")).read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "map_pdf_to_opacity")
    fn.returns = None
    for a in fn.args.args:
        a.annotation = None
    ns = {}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "encoder_spfsplatv2.py", "exec"), ns)
    map_pdf = ns["map_pdf_to_opacity"]
    for tag, (initial, final, warm_up, step) in {"e1": (0.0, 0.0, 1, 0), "warm": (0.0, 3.0, 1000, 500)}.items():
        fake = types.SimpleNamespace(cfg=types.SimpleNamespace(opacity_mapping=types.SimpleNamespace(initial=initial, final=final, warm_up=warm_up)))
        head = torch.randn(b, n, 83, generator=g)
        head[..., 1:] = raw.detach()
        head = head.requires_grad_()
        dens = head[..., 0].sigmoid()
        op = map_pdf(fake, dens, step)
        o2 = ad.forward(means, op, head[..., 1:])
        wo = torch.randn(b, n, generator=g)
        ((o2.opacities * wo).sum() + (o2.scales * ws).sum() + (o2.rotations * wr).sum() + (o2.harmonics * wh).sum()).backward()
        extra.update({f"head_{tag}": head.detach().numpy(), f"head_{tag}_cfg": np.array([initial, final, warm_up, step], dtype=np.float64),
                      f"head_{tag}_opacities": op.detach().numpy(), f"head_{tag}_wo": wo.numpy(), f"head_{tag}_d": head.grad.numpy()})
        print(f"head {tag}: opacity range {op.min().item():.4f}..{op.max().item():.4f}")
    np.savez_compressed(os.path.join(HERE, "adapter_ref.npz"), raw=raw.detach().numpy(), means=means.numpy(), opacities=opac.numpy(),
                        scales=out.scales.detach().numpy(), rotations=out.rotations.detach().numpy(),
                        harmonics=out.harmonics.detach().numpy(), ws=ws.numpy(), wr=wr.numpy(), wh=wh.numpy(),
                        d_raw=raw.grad.numpy(), sh_mask=ad.sh_mask.numpy(), **extra)
    print(f"adapter: scales max {out.scales.max().item():.3f}, d_raw norm {raw.grad.norm().item():.3f}")


def make_ply():
    """The reference's export_ply (src/model/ply_export.py:76-141), unmodified, on a seeded scene; the absent `plyfile`
    package is replaced by a recorder that keeps the element array the reference hands to PlyElement.describe (what
    plyfile would serialise: 17 'f4' properties per vertex)."""
    captured = {}
    m = types.ModuleType("plyfile")

    class PlyElement:
        @staticmethod
        def describe(elements, name):
            captured["elements"], captured["name"] = elements, name
            return ("element", name)

    class PlyData:
        def __init__(self, els):
            self.els = els

        def write(self, path):
            captured["path"] = str(path)
    m.PlyElement, m.PlyData = PlyElement, PlyData
    sys.modules["plyfile"] = m
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib
    mod = importlib.import_module("src.model.ply_export")
    g = torch.Generator().manual_seed(23)
    n = 1500
    means = torch.randn(n, 3, generator=g) * torch.tensor([2.0, 1.0, 4.0]) + torch.tensor([0.3, -0.2, 5.0])
    scales = torch.exp(torch.randn(n, 3, generator=g) * 0.7 - 4.0)
    rot = torch.randn(n, 4, generator=g)
    rot = rot / rot.norm(dim=-1, keepdim=True)
    rot[:4] = torch.tensor([[0.0, 0.0, 0.0, 1.0], [1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0]])   # all four pivots
    harm = torch.randn(n, 3, 25, generator=g)
    opac = torch.rand(n, generator=g)
    a = 0.4
    ext = torch.eye(4)
    ext[:3, :3] = torch.tensor([[math.cos(a), 0.0, math.sin(a)], [0.0, 1.0, 0.0], [-math.sin(a), 0.0, math.cos(a)]]) @ \
        torch.tensor([[1.0, 0.0, 0.0], [0.0, math.cos(0.2), -math.sin(0.2)], [0.0, math.sin(0.2), math.cos(0.2)]])
    ext[:3, 3] = torch.tensor([0.5, 0.1, -0.3])
    import tempfile
    from pathlib import Path
    mod.export_ply(ext, means, scales, rot, harm, opac, Path(tempfile.mkdtemp()) / "scene.ply")
    el = captured["elements"]
    names = list(el.dtype.names)
    rows = np.stack([el[k] for k in names], axis=1).astype(np.float32)
    assert names == mod.construct_list_of_attributes(0) and captured["name"] == "vertex" and rows.shape == (n, 17)
    np.savez_compressed(os.path.join(HERE, "ply_ref.npz"), extrinsics=ext.numpy(), means=means.numpy(), scales=scales.numpy(),
                        rotations=rot.numpy(), harmonics=harm.numpy(), opacities=opac.numpy(), rows=rows,
                        names=np.array(names))
    print(f"ply: {n} vertices, |xyz| 95% quantile {np.quantile(np.abs(rows[:, :3]), 0.95):.3f}")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        raise SystemExit("make_golden.py needs /root/reference (run it in the build container)")
    make_rope()
    make_decoder()
    make_orthographic()
    make_adapter()
    make_ply()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
