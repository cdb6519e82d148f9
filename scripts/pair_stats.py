"""CPU analysis of the blend kernels' lane efficiency on a bench workload (numpy, one view).

For every (Gaussian, tile) duplicate and each of the tile's eight 8x4 warp regions it counts
  hits        the duplicate's conservative alpha box (binning.cu::alpha_bbox) meets the region -> one iteration of
              the forward's inner loop with all 32 lanes evaluating alpha
  candidates  pixel centres of the region inside the box (what a pair-parallel forward over box pixels would evaluate)
  span        pixel centres inside the exact per-row ellipse span {alpha >= 1/255} widened by one pixel each side
  pairs       pixels with alpha >= 1/255 and power <= 0 (ignoring the transmittance stop): the useful lanes
so that design alternatives for blend_forward / blend_backward can be costed before they are written.

    python scripts/pair_stats.py [--workload c2p] [--view 0]

Test/analysis infrastructure only: imports the oracle's projection, never used by the product path.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2p")
    ap.add_argument("--view", type=int, default=0)
    args = ap.parse_args()
    import bench
    from oracle import raster_oracle as O
    from spfsplatv2_b200.camera import camera_setup

    v_cxt, h, w, b, desc = bench.WORKLOADS[args.workload]
    sc, _ = bench._make_inputs(args.workload, 0, pin=False)
    i = args.view
    view, proj, tanfov, scale = camera_setup(sc.extrinsics[i], sc.intrinsics[i], sc.near[i], sc.far[i], True)
    vw = O.View(h, w, float(tanfov[0, 0]), float(tanfov[0, 1]), torch.zeros(3), view[0].contiguous(),
                proj[0].contiguous(), 4, 1.0)
    with torch.no_grad():
        pre = O.preprocess(sc.means[i] * scale[0], sc.scales[i] * scale[0], sc.rotations[i], sc.opacities[i],
                           sc.harmonics[i].permute(0, 2, 1).contiguous(), None, vw)
    vis = pre["visible"].numpy()
    xy = pre["xy"].numpy()[vis].astype(np.float64)
    con = pre["conic"].numpy()[vis].astype(np.float64)
    op = pre["opacity"].numpy()[vis].astype(np.float64)
    rect = pre["rect"].numpy()[vis].astype(np.int64)        # tile rect [x0,y0,x1,y1)
    P = int(vis.sum())

    tau = 2.0 * np.log(np.maximum(255.0 * op, 1e-30))
    det = con[:, 0] * con[:, 2] - con[:, 1] ** 2
    never = tau < 0
    ex = np.sqrt(np.maximum(tau, 0) * con[:, 2] / det) * 1.0001 + 0.01
    ey = np.sqrt(np.maximum(tau, 0) * con[:, 0] / det) * 1.0001 + 0.01

    n_dup = hits = cand = span = pairs = zero_hits = 0
    per_tile = np.zeros(((h + 15) // 16) * ((w + 15) // 16), dtype=np.int64)
    gx = (w + 15) // 16
    for g in range(P):
        x0t, y0t, x1t, y1t = rect[g]
        n_dup += (x1t - x0t) * (y1t - y0t)
        for ty in range(y0t, y1t):
            for tx in range(x0t, x1t):
                per_tile[ty * gx + tx] += 1
        if never[g]:
            continue
        bx0, bx1, by0, by1 = xy[g, 0] - ex[g], xy[g, 0] + ex[g], xy[g, 1] - ey[g], xy[g, 1] + ey[g]
        # pixel centres (integers) inside the box, clipped to the image and to the Gaussian's own tile rect
        px0, px1 = max(int(np.ceil(bx0)), x0t * 16, 0), min(int(np.floor(bx1)), x1t * 16 - 1, w - 1)
        py0, py1 = max(int(np.ceil(by0)), y0t * 16, 0), min(int(np.floor(by1)), y1t * 16 - 1, h - 1)
        if px1 < px0 or py1 < py0:
            continue
        xs, ys = np.meshgrid(np.arange(px0, px1 + 1), np.arange(py0, py1 + 1))
        dx, dy = xy[g, 0] - xs, xy[g, 1] - ys
        power = -0.5 * (con[g, 0] * dx * dx + con[g, 2] * dy * dy) - con[g, 1] * dx * dy
        alpha = np.minimum(0.99, op[g] * np.exp(power))
        okp = (power <= 0) & (alpha >= 1.0 / 255.0)
        # exact row span widened by one pixel each side: per row, columns between first-1 and last+1 contributing
        spanp = np.zeros_like(okp)
        for r in range(okp.shape[0]):
            cols = np.nonzero(okp[r])[0]
            if cols.size:
                spanp[r, max(cols[0] - 1, 0):cols[-1] + 2] = True
        region = (ys // 4) * 1024 + (xs // 8)               # id of the 8x4 warp region
        for rid in np.unique(region):
            m = region == rid
            hits += 1
            cand += int(m.sum())
            span += int(spanp[m].sum())
            k = int(okp[m].sum())
            pairs += k
            zero_hits += (k == 0)
    T = per_tile.size
    res = {
        "workload": args.workload, "view": i, "visible_gaussians": P, "duplicates": int(n_dup),
        "records_per_tile": round(float(per_tile.mean()), 1),
        "hits_per_warp_tile": round(hits / (8.0 * T), 1),
        "box_pixels_per_hit": round(cand / max(hits, 1), 2),
        "span_pixels_per_hit": round(span / max(hits, 1), 2),
        "contributing_pixels_per_hit": round(pairs / max(hits, 1), 2),
        "zero_contribution_hits": round(zero_hits / max(hits, 1), 3),
        "pairs_per_warp_tile": round(pairs / (8.0 * T), 1),
    }
    print(json.dumps(res))


if __name__ == "__main__":
    main()
