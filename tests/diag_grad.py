"""Diagnostic (not a test): gradient error of the CUDA path and of the fp32 oracle against a float64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spfsplatv2_b200.synthetic import make_batch
from tests.util import oracle_views, rel_err

def run(regime, h, w, grid, b, v, seed=11):
    sc = make_batch(b, seed=seed, v_cxt=1, h=h, w=w, grid=grid, regime=regime, n_target=v, with_cov=True)
    bg = (0.2, 0.1, 0.4)
    torch.manual_seed(0)
    wc = torch.randn(b * v, 3, h, w)
    wd = 0.05 * torch.randn(b * v, 1, h, w)
    res = {}
    for name, dt in (("o32", torch.float32), ("o64", torch.float64)):
        ref, leaves = oracle_views(sc, bg=bg, requires_grad=True, dtype=dt)
        loss = sum((r["color"] * wc[i].to(dt)).sum() + (r["depth"] * sc.near.reshape(-1)[i] * wd[i].to(dt)).sum() for i, r in enumerate(ref))
        loss.backward()
        res[name] = {k: t.grad for k, t in leaves.items()}
        res[name + "_nc"] = [r["n_contrib"] for r in ref]
    same = all(torch.equal(a, c) for a, c in zip(res["o32_nc"], res["o64_nc"]))
    from spfsplatv2_b200.decoder import DecoderSplattingCUDA, DecoderSplattingCUDACfg, Gaussians
    d = torch.device("cuda:0")
    dec = DecoderSplattingCUDA(DecoderSplattingCUDACfg("splatting_cuda", list(bg), True, True, True)).to(d)
    t = {k: getattr(sc, k).to(d).requires_grad_() for k in ("means", "rotations", "scales", "harmonics", "opacities")}
    ext = sc.extrinsics.to(d).requires_grad_()
    g = Gaussians(t["means"], sc.covariances.to(d), t["rotations"], t["scales"], t["harmonics"], t["opacities"])
    out = dec(g, ext, sc.intrinsics.to(d), sc.near.to(d), sc.far.to(d), sc.image_shape)
    l2 = (out.color.reshape(b * v, 3, h, w) * wc.to(d)).sum() + (out.depth.reshape(b * v, 1, h, w) * wd.to(d)).sum()
    l2.backward()
    ours = {k: t[k].grad.cpu() for k in t}; ours["extrinsics"] = ext.grad.cpu()
    print(f"--- {regime} {h}x{w} grid={grid} b={b} v={v}  n_contrib o32==o64: {same}")
    for k in ours:
        print(f"  {k:10s} ours-vs-o32 {rel_err(ours[k], res['o32'][k]):.2e}  ours-vs-o64 {rel_err(ours[k], res['o64'][k]):.2e}  o32-vs-o64 {rel_err(res['o32'][k], res['o64'][k]):.2e}")
    k = "means"
    e = (ours[k].double() - res["o64"][k]).abs().flatten()
    top = e.topk(5)
    print("  worst means entries:", [(int(i) // 3, f"{ours[k].flatten()[i]:.4e}", f"{res['o64'][k].flatten()[i]:.4e}", f"{res['o32'][k].flatten()[i]:.4e}") for i in top.indices])

if __name__ == "__main__":
    run("trained", 96, 96, (48, 48), 1, 1)
    run("trained", 64, 48, (24, 24), 2, 3)
    run("init", 64, 64, (32, 32), 1, 1)
    run("trained", 128, 128, (64, 64), 1, 1, seed=3)
