"""Host-side profile of the level-0 integration path: the reference's per-view loop over the diff_gauss_pose drop-in."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import WORKLOADS, _make_inputs
from spfsplatv2_b200 import diff_gauss_pose as shim
from spfsplatv2_b200.loss import mse_loss
from tests.ref_probe import reference_render_loop

v_cxt, h, w, b, _ = WORKLOADS["c2p"]
dev = torch.device("cuda:0")
sc, host = _make_inputs("c2p", 0, pin=False)
d = {k: v.to(dev) for k, v in host.items()}
names = ("means", "scales", "rotations", "opacities", "harmonics", "extrinsics")
bg = torch.zeros(b, 3, device=dev)


def step(split=None):
    t0 = time.perf_counter()
    leaves = {k: d[k].detach().requires_grad_() for k in names}
    scd = sc.__class__(leaves["means"], None, leaves["rotations"], leaves["scales"], leaves["harmonics"], leaves["opacities"],
                       leaves["extrinsics"], d["intrinsics"], d["near"], d["far"], (h, w))
    color, _ = reference_render_loop(shim, scd, bg, leaves=leaves)
    t1 = time.perf_counter()
    loss = mse_loss(color, d["gt"][:, 0])
    loss.backward()
    t2 = time.perf_counter()
    if split is not None:
        split[0] += t1 - t0
        split[1] += t2 - t1
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
split = [0.0, 0.0]
t0 = time.perf_counter()
for _ in range(10):
    step(split)
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t0) * 100, "forward loop", split[0] * 100, "loss+backward", split[1] * 100)
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
