"""Host-side (CPU) time per step of the decoder path, split by phase; GPU kept busy, no syncs inside."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import _make_inputs, WORKLOADS
from spfsplatv2_b200.decoder import DecoderSplattingCUDA, DecoderSplattingCUDACfg, Gaussians
from spfsplatv2_b200.loss import mse_loss

wl = sys.argv[1] if len(sys.argv) > 1 else "c2p"
v_cxt, h, w, b, _ = WORKLOADS[wl]
dev = torch.device("cuda:0")
sc, host = _make_inputs(wl, 0, pin=False)
P = sc.means.shape[1]
dec = DecoderSplattingCUDA(DecoderSplattingCUDACfg("splatting_cuda", [0.0, 0.0, 0.0], True, True, True)).to(dev)
d = {k: v.to(dev) for k, v in host.items()}
cov = torch.zeros(1, 1, 3, 3, device=dev).expand(b, P, 3, 3)
acc = {"leaves": 0.0, "forward": 0.0, "loss": 0.0, "backward": 0.0}
N = 60
for it in range(N + 10):
    if it == 10:
        torch.cuda.synchronize(); acc = {k: 0.0 for k in acc}; t_start = time.perf_counter()
    t0 = time.perf_counter()
    leaves = {k: d[k].detach().requires_grad_() for k in ("means", "rotations", "scales", "harmonics", "opacities")}
    ext = d["extrinsics"].detach().requires_grad_()
    g = Gaussians(leaves["means"], cov, leaves["rotations"], leaves["scales"], leaves["harmonics"], leaves["opacities"])
    t1 = time.perf_counter()
    out = dec(g, ext, d["intrinsics"], d["near"], d["far"], (h, w))
    t2 = time.perf_counter()
    loss = mse_loss(out.color, d["gt"])
    t3 = time.perf_counter()
    loss.backward()
    t4 = time.perf_counter()
    acc["leaves"] += t1 - t0; acc["forward"] += t2 - t1; acc["loss"] += t3 - t2; acc["backward"] += t4 - t3
torch.cuda.synchronize()
tot = time.perf_counter() - t_start
print(wl, "wall ms/step", 1e3 * tot / N, {k: round(1e3 * v / N, 3) for k, v in acc.items()})
# finer: inside forward
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for it in range(30):
    leaves = {k: d[k].detach().requires_grad_() for k in ("means", "rotations", "scales", "harmonics", "opacities")}
    ext = d["extrinsics"].detach().requires_grad_()
    g = Gaussians(leaves["means"], cov, leaves["rotations"], leaves["scales"], leaves["harmonics"], leaves["opacities"])
    out = dec(g, ext, d["intrinsics"], d["near"], d["far"], (h, w))
    loss = mse_loss(out.color, d["gt"])
    loss.backward()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(45)
