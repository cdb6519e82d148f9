"""Shared definitions of the FULL-SIZE oracle fixtures (BASELINE.json configs 2-4 and the headline "65k" scene):
which scenes, which seeded loss, and the compact summaries that are committed instead of the raw tensors.
Used by tests/golden/make_golden_fullsize.py (writes tests/golden/fullsize_<name>.npz from the CPU oracle) and by
tests/test_fullsize_gpu.py / tests/test_golden_cpu.py (read them).  Test infrastructure only.

Compact summaries
  * images: colour in full at 256x256, 2x decimated at 512x512, plus fp64 per-tile sums of every pixel;
    PSNR against a seeded pseudo ground truth (src/evaluation/metrics.py:12-19 definition);
  * indices: radii / tiles_touched / n_contrib / tile ranges in full (small integer arrays, compressed),
    sha256 + per-tile (sum, xor) checksums of the sorted 64-bit keys and of the point list;
  * gradients: the full camera-pose gradient; for each Gaussian tensor its fp64 L2 norm, every `stride`-th row,
    and N_PROJ sign projections <g, r_k> (r_k in {-1,+1}^n from an integer hash, identical on any device), from
    which the L2 error of a candidate gradient is estimated: E[<d, r>^2] = |d|^2.
"""
from __future__ import annotations

import hashlib

import numpy as np
import torch

# name: dict(v_cxt, h, w, regime, seed, bg)
CONFIGS = {
    "c2p": dict(v_cxt=1, h=256, w=256, regime="init", seed=0, bg=(0.0, 0.0, 0.0)),        # headline: P = 65 536
    "c2": dict(v_cxt=2, h=256, w=256, regime="init", seed=1, bg=(0.0, 0.0, 0.0)),         # re10k 2-view: P = 131 072
    "c3": dict(v_cxt=10, h=256, w=256, regime="init", seed=2, bg=(0.1, 0.2, 0.3)),        # re10k 10-view: P = 655 360
    "c4i": dict(v_cxt=2, h=512, w=512, regime="init", seed=3, bg=(0.0, 0.0, 0.0)),        # acid 2-view 512^2: P = 524 288
    "c4t": dict(v_cxt=2, h=512, w=512, regime="trained", seed=4, bg=(0.3, 0.2, 0.1)),     # same, trained-like splat sizes
}
N_PROJ = 32
GRAD_NAMES = ("means", "scales", "rotations", "opacities", "harmonics")


def scene_of(name: str):
    from spfsplatv2_b200.synthetic import make_scene
    c = CONFIGS[name]
    return make_scene(seed=c["seed"], v_cxt=c["v_cxt"], h=c["h"], w=c["w"], regime=c["regime"], n_target=1)


def loss_weights(name: str):
    """Seeded upstream weights: loss = sum(color * wc) + sum(depth_scaled * wd)."""
    c = CONFIGS[name]
    g = torch.Generator().manual_seed(1000 + c["seed"])
    wc = torch.randn(1, 3, c["h"], c["w"], generator=g)
    wd = 0.05 * torch.randn(1, 1, c["h"], c["w"], generator=g)
    return wc, wd


def pseudo_gt(name: str):
    c = CONFIGS[name]
    return torch.rand(1, 3, c["h"], c["w"], generator=torch.Generator().manual_seed(2000 + c["seed"]))


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


def inputs_digest(sc) -> str:
    return sha(torch.cat([x.reshape(-1) for x in (sc.means, sc.scales, sc.rotations, sc.opacities, sc.harmonics,
                                                  sc.extrinsics, sc.intrinsics, sc.near, sc.far)]))


def sign_vector(n: int, k: int, device) -> torch.Tensor:
    """r_k in {-1,+1}^n as float64, from a 32-bit integer hash of (index, k): same values on CPU and GPU."""
    i = torch.arange(n, dtype=torch.int64, device=device)
    m = 0xFFFFFFFF
    h = (i * 2654435761 + (k + 1) * 40503) & m
    h = h ^ (h >> 15)
    h = (h * 2246822519) & m
    h = h ^ (h >> 13)
    h = (h * 3266489917) & m
    h = h ^ (h >> 16)
    return ((h >> 7) & 1).to(torch.float64) * 2.0 - 1.0


def projections(g: torch.Tensor) -> torch.Tensor:
    flat = g.detach().reshape(-1).to(torch.float64)
    return torch.stack([(flat * sign_vector(flat.numel(), k, flat.device)).sum() for k in range(N_PROJ)])


def stride_of(n_rows: int) -> int:
    return max(1, n_rows // 1024)


def tile_sums(img: torch.Tensor) -> torch.Tensor:
    """[C,H,W] -> fp64 [C, ceil(H/16), ceil(W/16)] sums over 16x16 tiles."""
    C, H, W = img.shape
    gy, gx = (H + 15) // 16, (W + 15) // 16
    pad = torch.zeros(C, gy * 16, gx * 16, dtype=torch.float64, device=img.device)
    pad[:, :H, :W] = img.to(torch.float64)
    return pad.view(C, gy, 16, gx, 16).sum(dim=(2, 4))


def key_tile_checksums(keys: torch.Tensor, point_list: torch.Tensor, ranges: torch.Tensor):
    """Per-tile (sum of depth bits, sum of Gaussian ids weighted by list position mod 251) -- order-sensitive."""
    T = ranges.shape[0]
    n = keys.numel()
    pos = torch.arange(n, dtype=torch.int64, device=keys.device)
    tile = (keys >> 32).to(torch.int64)
    wgt = (pos % 251) + 1
    a = torch.zeros(T, dtype=torch.int64, device=keys.device).index_add_(0, tile, (keys & 0xFFFFFFFF) * wgt)
    b = torch.zeros(T, dtype=torch.int64, device=keys.device).index_add_(0, tile, point_list.to(torch.int64) * wgt)
    return a, b


def small_int(t: torch.Tensor) -> np.ndarray:
    t = t.detach().cpu()
    mx = int(t.max()) if t.numel() else 0
    return t.numpy().astype(np.int16 if mx < 32768 else np.int32)
