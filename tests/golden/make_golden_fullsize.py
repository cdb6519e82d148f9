#!/usr/bin/env python
"""Writes tests/golden/fullsize_<name>.npz: the CPU oracle (oracle/raster_oracle.py, forward + autograd backward) on ONE
view of each full-size BASELINE.json configuration -- the headline 65k scene (c2p), re10k 2-view (c2), re10k 10-view
(c3) and acid 2-view 512x512 in both scale regimes (c4i, c4t).  Compact summaries only (tests/golden/fullsize.py).

    python tests/golden/make_golden_fullsize.py [names...]        # ~1-10 s of oracle per view, minutes for c3/c4t

The rasterizer oracle is "parity unpinned" (no reference build or vectors exist for diff_gauss_pose); these
fixtures pin the CUDA path to the oracle at the sizes the benchmark runs, they do not pin the oracle to the reference.
Test infrastructure only.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import raster_oracle as O            # noqa: E402
from tests.golden import fullsize as F           # noqa: E402
from tests.util import oracle_views              # noqa: E402


def make(name: str):
    cfg = F.CONFIGS[name]
    h, w = cfg["h"], cfg["w"]
    sc = F.scene_of(name)
    wc, wd = F.loss_weights(name)
    t0 = time.perf_counter()
    res, leaves = oracle_views(sc, bg=cfg["bg"], requires_grad=True)
    r = res[0]
    near = sc.near.reshape(-1)[0]
    depth_scaled = r["depth"] * near
    loss = (r["color"] * wc[0]).sum() + (depth_scaled * wd[0]).sum()
    t1 = time.perf_counter()
    loss.backward()
    t2 = time.perf_counter()
    color = r["color"].detach()
    gt = F.pseudo_gt(name)
    out = dict(
        inputs_sha=np.array(F.inputs_digest(sc)),
        loss=np.float64(loss.item()),
        psnr=np.float64(O.compute_psnr(gt, color[None]).item()),
        color_tile_sums=F.tile_sums(color).numpy(),
        depth_tile_sums=F.tile_sums(depth_scaled.detach()).numpy(),
        alpha_tile_sums=F.tile_sums(r["alpha"].detach()).numpy(),
        n_contrib=F.small_int(r["n_contrib"]),
        radii=F.small_int(r["pre"]["radius"]),
        tiles_touched=F.small_int(r["pre"]["tiles_touched"]),
        ranges=r["ranges"].numpy(),
        n_dups=np.int64(r["keys"].numel()),
        keys_sha=np.array(F.sha(r["keys"])),
        point_list_sha=np.array(F.sha(r["point_list"])),
        grad_extrinsics=leaves["extrinsics"].grad.numpy(),
        oracle_seconds=np.array([t1 - t0, t2 - t1]),
    )
    dec = 1 if h * w <= 256 * 256 else 2
    out["color"] = color[:, ::dec, ::dec].contiguous().numpy()
    out["depth"] = depth_scaled.detach()[:, ::2, ::2].contiguous().numpy()
    out["decimation"] = np.array([dec, 2])
    ka, kb = F.key_tile_checksums(r["keys"], r["point_list"], r["ranges"])
    out["key_tile_sums"], out["point_tile_sums"] = ka.numpy(), kb.numpy()
    for nme in F.GRAD_NAMES:
        g = leaves[nme].grad[0]                                  # [P, ...]
        st = F.stride_of(g.shape[0])
        out[f"gnorm_{nme}"] = np.float64(g.double().norm().item())
        out[f"gproj_{nme}"] = F.projections(g).numpy()
        out[f"grows_{nme}"] = g[::st].contiguous().numpy()
        out[f"gsum_{nme}"] = g.double().reshape(g.shape[0], -1).sum(0).numpy()
    path = os.path.join(HERE, f"fullsize_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: P={sc.means.shape[1]} N={int(out['n_dups'])} psnr={float(out['psnr']):.4f} loss={loss.item():.4f} "
          f"oracle fwd {t1 - t0:.1f}s bwd {t2 - t1:.1f}s -> {os.path.getsize(path) / 1e6:.2f} MB", flush=True)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    for nm in (sys.argv[1:] or list(F.CONFIGS)):
        make(nm)
