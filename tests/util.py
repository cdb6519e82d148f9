"""Shared helpers for the parity tests: run the CPU oracle on a synthetic Scene."""
import torch

from oracle import raster_oracle as O
from spfsplatv2_b200.camera import camera_setup


def oracle_views(sc, scale_invariant=True, bg=(0.0, 0.0, 0.0), requires_grad=False, use_sh=True, dtype=torch.float32,
                 **render_kw):
    """Render every (scene, view) of ``sc`` with the oracle.  Returns (list of per-view result
    dicts in (b v) order, leaves dict) -- leaves are CPU tensors with requires_grad for autograd."""
    b, v = sc.extrinsics.shape[:2]
    h, w = sc.image_shape
    leaves = dict(means=sc.means.clone(), scales=sc.scales.clone(), rotations=sc.rotations.clone(),
                  opacities=sc.opacities.clone(), harmonics=sc.harmonics.clone(), extrinsics=sc.extrinsics.clone())
    leaves = {k: t.to(dtype) for k, t in leaves.items()}     # float64 = truth check for gradient tests only
    if requires_grad:
        for t in leaves.values():
            t.requires_grad_()
    ext = leaves["extrinsics"].reshape(b * v, 4, 4)
    view, proj, tanfov, scale = camera_setup(ext, sc.intrinsics.reshape(b * v, 3, 3).to(dtype), sc.near.reshape(-1).to(dtype),
                                             sc.far.reshape(-1).to(dtype), scale_invariant)
    out = []
    for i in range(b * v):
        s = i // v
        vw = O.View(h, w, float(tanfov[i, 0]), float(tanfov[i, 1]), torch.tensor(bg, dtype=dtype),
                    view[i].contiguous(), proj[i].contiguous(), int(sc.harmonics.shape[-1] ** 0.5 + 0.5) - 1, 1.0)
        shs = leaves["harmonics"][s].permute(0, 2, 1).contiguous()
        res = O.render(leaves["means"][s] * scale[i], leaves["scales"][s] * scale[i], leaves["rotations"][s],
                       leaves["opacities"][s], shs if use_sh else None, None if use_sh else shs[:, 0, :], vw, **render_kw)
        out.append(res)
    return out, leaves


def rel_err(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()
