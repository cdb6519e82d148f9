"""Probe for the REAL reference rasterizer and drive it the way the reference does.  Test / benchmark infrastructure.

SPFSplatV2's rasterizer is the external CUDA package ``diff_gauss_pose`` (requirements.txt:87, imported at
src/model/decoder/cuda_splatting.py:5).  It is not vendored, not in this image and cannot be fetched; if a driver or a
user provides it (site-packages, or unpacked under ``baseline/_ref/``) the A/B legs switch on by themselves:
``bench.py --impl reference`` times it on the GPU (``cpu_baseline.kind = "reference-cuda"``) and
``tests/test_ref_ab_gpu.py`` compares images and gradients with ours.  Otherwise both say why not.

``reference_render_loop`` restates, per view, what cuda_splatting.py:96-143 does with the package: one settings record
and one rasterizer call per view, Python-float tanfov, transposed matrices, SH as [P, K, 3], opacities as [P, 1], a zero
``means2D`` tensor; the host glue in front of it (scale invariance, fov, projection matrix, inverse) is
spfsplatv2_b200.camera.camera_setup, which tests/test_golden_cpu.py pins bit for bit against the reference's own code.
"""
from __future__ import annotations

import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference_rasterizer():
    """(module, None) if a genuine diff_gauss_pose is importable, else (None, reason).  Our own drop-in of the same
    name (spfsplatv2_b200.diff_gauss_pose, installable into sys.modules by install_shims) never counts."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    added = False
    if os.path.isdir(ref_dir) and ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
        added = True
    try:
        cached = sys.modules.get("diff_gauss_pose")
        if cached is not None and "spfsplatv2_b200" in (getattr(cached, "__file__", "") or ""):
            del sys.modules["diff_gauss_pose"]          # our shim was installed under the reference's name: look past it
        try:
            mod = importlib.import_module("diff_gauss_pose")
        except Exception as exc:     # ImportError, or a binary built for another torch / arch
            return None, f"diff_gauss_pose not importable ({type(exc).__name__}: {exc}); not vendored under /root/reference, no network"
        f = getattr(mod, "__file__", "") or ""
        if "spfsplatv2_b200" in f:
            return None, "only this repo's own drop-in answers to the name diff_gauss_pose"
        if not (hasattr(mod, "GaussianRasterizationSettings") and hasattr(mod, "GaussianRasterizer")):
            return None, f"{f}: no GaussianRasterizationSettings / GaussianRasterizer"
        return mod, None
    finally:
        if added and "diff_gauss_pose" not in sys.modules:
            sys.path.remove(ref_dir)


def reference_render_loop(mod, sc, background, use_sh: bool = True, enable_cov_grad: bool = True,
                          enable_sh_grad: bool = True, leaves=None):
    """All (scene, view) pairs of ``sc`` through ``mod`` (the real package, or anything with its interface), one call
    per view as the reference loops.  Returns (color [B,3,H,W], depth [B,1,H,W] scaled back by near)."""
    from spfsplatv2_b200.camera import camera_setup
    dev = sc.means.device
    b, v = sc.extrinsics.shape[:2]
    h, w = sc.image_shape
    t = leaves if leaves is not None else {k: getattr(sc, k) for k in ("means", "scales", "rotations", "opacities", "harmonics", "extrinsics")}
    view, proj, tanfov, scale = camera_setup(t["extrinsics"].reshape(b * v, 4, 4), sc.intrinsics.reshape(b * v, 3, 3),
                                             sc.near.reshape(-1), sc.far.reshape(-1), True)
    K = sc.harmonics.shape[-1]
    degree = int(K ** 0.5 + 0.5) - 1
    tan_host = tanfov.detach().cpu()
    images, depths = [], []
    for i in range(b * v):
        s = i // v
        settings = mod.GaussianRasterizationSettings(
            image_height=h, image_width=w, tanfovx=float(tan_host[i, 0]), tanfovy=float(tan_host[i, 1]),
            bg=background[i], scale_modifier=1.0, projmatrix=proj[i], sh_degree=degree, prefiltered=False, debug=False,
            enable_cov_grad=enable_cov_grad, enable_sh_grad=enable_sh_grad)
        means = t["means"][s] * scale[i]
        shs = t["harmonics"][s].permute(0, 2, 1).contiguous()
        out = mod.GaussianRasterizer(settings)(
            means3D=means, means2D=torch.zeros_like(means, requires_grad=True),
            shs=shs if use_sh else None, colors_precomp=None if use_sh else shs[:, 0, :],
            opacities=t["opacities"][s][:, None], scales=t["scales"][s] * scale[i], rotations=t["rotations"][s],
            viewmatrix=view[i])
        images.append(out[0])
        depths.append(out[1])
    color = torch.stack(images)
    depth = torch.stack(depths).reshape(b * v, 1, h, w) * sc.near.reshape(-1).to(dev)[:, None, None, None]
    return color, depth


def reference_tree():
    """Path of an unmodified copy of the reference's python tree under baseline/_ref (placed there by whoever runs the
    test: `cp -r /root/reference/src baseline/_ref/src`; never shipped with the repo), or None."""
    d = os.path.join(ROOT, "baseline", "_ref")
    return d if os.path.isfile(os.path.join(d, "src", "model", "decoder", "cuda_splatting.py")) else None


def stub_absent_third_party_modules():
    """The reference's decoder package imports (transitively) a dozen third-party modules this image does not have
    (dacite, lightning, skvideo, ...); none of them is touched on the decoder path.  Absent ones are
    replaced by permissive placeholders, present ones are left alone -- as tests/golden/make_golden.py does."""
    import types
    import torchvision  # noqa: F401  (the real one first)

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
