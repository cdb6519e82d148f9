"""Soak test of the host-side sizing machinery: a few hundred training / inference calls over changing scenes, regimes,
view counts and image sizes (capacity growth, deferred duplicate-count checks, pair-log capacity feedback, overflow
re-runs, raw-head and standard inputs interleaved), every result checked for finiteness and every 10th against a
sync_count=True render of the same inputs.  python scripts/stress.py [seconds]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from spfsplatv2_b200.decoder import DecoderSplattingCUDA, DecoderSplattingCUDACfg, Gaussians
from spfsplatv2_b200.rasterizer import DuplicateCapacityError
from spfsplatv2_b200.synthetic import make_batch

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
dev = torch.device("cuda:0")
dec = DecoderSplattingCUDA(DecoderSplattingCUDACfg("splatting_cuda", [0.1, 0.0, 0.2], True, True, True)).to(dev)
g = torch.Generator().manual_seed(0)
shapes = [(64, 64, (32, 32)), (96, 80, (48, 40)), (128, 128, (64, 64)), (256, 256, None), (112, 112, (56, 56))]
t0 = time.time()
it = n_train = n_infer = n_retry = n_head = 0
while time.time() - t0 < budget:
    h, w, grid = shapes[int(torch.randint(len(shapes), (1,), generator=g))]
    regime = "trained" if torch.rand(1, generator=g).item() < 0.5 else "init"
    b = int(torch.randint(1, 4, (1,), generator=g))
    v = int(torch.randint(1, 3, (1,), generator=g))
    sc = make_batch(b, seed=it * 7 + 1, v_cxt=1, h=h, w=w, grid=grid, regime=regime, n_target=v).to(dev)
    if regime == "trained" and torch.rand(1, generator=g).item() < 0.3:
        sc.scales = sc.scales * 3.0          # a sudden jump in splat size: duplicate count far above the high-water mark
    train = torch.rand(1, generator=g).item() < 0.7
    leaves = {k: getattr(sc, k).clone().requires_grad_(train) for k in ("means", "rotations", "scales", "harmonics", "opacities")}
    ext = sc.extrinsics.clone().requires_grad_(train)
    G = Gaussians(leaves["means"], sc.covariances, leaves["rotations"], leaves["scales"], leaves["harmonics"], leaves["opacities"])
    for attempt in range(3):
        try:
            with torch.set_grad_enabled(train):
                out = dec(G, ext, sc.intrinsics, sc.near, sc.far, sc.image_shape)
                if train:
                    (out.color.square().mean() + 0.01 * out.depth.mean()).backward()
            break
        except DuplicateCapacityError:
            n_retry += 1
            for t in list(leaves.values()) + [ext]:
                t.grad = None
    else:
        raise SystemExit("three DuplicateCapacityErrors in a row")
    assert torch.isfinite(out.color).all() and torch.isfinite(out.depth).all(), (it, "non-finite image")
    if train:
        for k, t in leaves.items():
            assert t.grad is not None and torch.isfinite(t.grad).all(), (it, k)
        assert torch.isfinite(ext.grad).all()
        n_train += 1
    else:
        n_infer += 1
    if it % 10 == 0:       # exact-count inference render of the same inputs must agree with what we got
        with torch.no_grad():
            ref = dec(G, ext, sc.intrinsics, sc.near, sc.far, sc.image_shape)
        assert torch.equal(ref.color, out.color.detach()), (it, "image differs from the exact-count render")
    if it % 7 == 0 and sc.means.shape[1] % 4 == 0:      # raw-head entry on the same cameras
        head = torch.randn(b, sc.means.shape[1], 83, generator=g).to(dev)
        head[..., 1:4] = head[..., 1:4] * 2 + 3
        head.requires_grad_(True)
        for attempt in range(3):
            try:
                o = dec.forward_head(sc.means, head, sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape, sh_degree=4)
                o.color.mean().backward()
                break
            except DuplicateCapacityError:
                n_retry += 1
                head.grad = None
        assert torch.isfinite(o.color).all() and torch.isfinite(head.grad).all(), (it, "raw head")
        n_head += 1
    it += 1
torch.cuda.synchronize()
print(f"stress ok: {it} iterations in {time.time() - t0:.0f} s ({n_train} training, {n_infer} inference, {n_head} raw-head, "
      f"{n_retry} capacity retries), peak memory {torch.cuda.max_memory_allocated() / 2**20:.0f} MiB")
