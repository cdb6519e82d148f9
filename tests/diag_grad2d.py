"""Diagnostic (not a test): per-Gaussian 2-D gradients out of blend_backward vs the oracle's autograd."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import raster_oracle as O
from spfsplatv2_b200 import _lib as L
from spfsplatv2_b200.camera import camera_setup
from spfsplatv2_b200.rasterizer import RasterSettings, forward_with_state, _ptr, _stream
from spfsplatv2_b200.synthetic import make_batch

h = w = 96
sc = make_batch(1, seed=11, v_cxt=1, h=h, w=w, grid=(48, 48), regime="trained", n_target=1, with_cov=True)
bg = (0.2, 0.1, 0.4)
torch.manual_seed(0)
wc = torch.randn(1, 3, h, w); wd = 0.05 * torch.randn(1, 1, h, w)
view, proj, tanfov, scale = camera_setup(sc.extrinsics[0], sc.intrinsics[0], sc.near[0], sc.far[0], True)
means = (sc.means[0] * scale[0]).contiguous(); scales = (sc.scales[0] * scale[0]).contiguous()
quats = sc.rotations[0].contiguous(); shs = sc.harmonics[0].permute(0, 2, 1).contiguous(); opac = sc.opacities[0].clone()
vw = O.View(h, w, float(tanfov[0, 0]), float(tanfov[0, 1]), torch.tensor(bg), view[0].contiguous(), proj[0].contiguous(), 4, 1.0)
def oracle2d(dt):
    vw2 = O.View(h, w, vw.tanfovx, vw.tanfovy, vw.bg.to(dt), vw.viewmatrix.to(dt), vw.projmatrix.to(dt), 4, 1.0)
    pre = O.preprocess(means.to(dt), scales.to(dt), quats.to(dt), opac.to(dt).requires_grad_(), shs.to(dt), None, vw2)
    pre["depth"] = pre["depth"] * 1.0
    pre["opacity"] = pre["opacity"] * 1.0
    for k in ("xy", "conic", "rgb", "depth", "opacity"):
        pre[k].requires_grad_() if not pre[k].requires_grad else None
        pre[k].retain_grad()
    keys, pl, ranges = O.bin_and_sort(pre, vw2)
    color, depth, alpha, fT, nc = O.blend(pre, pl, ranges, vw2)
    ((color * wc[0].to(dt)).sum() + (depth * float(sc.near.reshape(-1)[0]) * wd[0].to(dt)).sum()).backward()
    g = torch.zeros(means.shape[0], 10, dtype=dt)
    g[:, 0:2] = pre["xy"].grad; g[:, 2:5] = pre["conic"].grad; g[:, 5] = pre["opacity"].grad
    g[:, 6:9] = pre["rgb"].grad; g[:, 9] = pre["depth"].grad
    return g, pre, pl, ranges, nc
g32, pre, pl, ranges, nc = oracle2d(torch.float32)
g64 = oracle2d(torch.float64)[0]

d = torch.device("cuda:0")
s = RasterSettings(h, w, 4, 1.0, 1, sh_layout_ck=False)
color, depth, alpha, radii, st = forward_with_state(s, means[None].to(d), scales[None].to(d), quats[None].to(d), opac[None].to(d),
                                                    shs[None].to(d), None, view.to(d), proj.to(d), tanfov.to(d),
                                                    torch.tensor([bg], device=d), None)
P = means.shape[0]
f32 = dict(dtype=torch.float32, device=d)
gc = wc.to(d).contiguous(); gd = (wd * float(sc.near.reshape(-1)[0])).to(d).contiguous()
gout = L.SpfRasterGradOut(_ptr(gc), _ptr(gd), None)
dup = torch.zeros(st.n_dups, 12, **f32)
bufs = [torch.empty(1, (P + 127) // 128, 16, **f32), torch.empty(1, P, 3, **f32), torch.empty(1, P, 3, **f32), torch.empty(1, P, 4, **f32),
        torch.empty(1, P, **f32), torch.empty(1, P, 25, 3, **f32)]
dview = torch.empty(1, 16, **f32)
gin = L.SpfRasterGradIn(_ptr(dup), *[_ptr(b) for b in bufs], None, _ptr(dview), None)
L.check(L.lib().spf_raster_backward_stages(C.byref(st.desc), C.byref(st.cin), C.byref(st.cstate), C.byref(gout), C.byref(gin), 1, _stream(d)), "bwd")
torch.cuda.synchronize()
dup = dup.cpu(); off = st.tensors["dup_offset"][0].cpu(); tt = st.tensors["tiles_touched"][0].cpu()
ours = torch.zeros(P, 10)
for g in range(P):
    if tt[g] > 0:
        ours[g] = dup[off[g]:off[g] + tt[g], :10].sum(0)
names = ["dpx", "dpy", "dconx", "dcony", "dconz", "dopac", "dr", "dg", "db", "ddepth"]
print("n_contrib equal:", torch.equal(st.tensors["n_contrib"][0].cpu(), nc))
for k, n in enumerate(names):
    e1 = (ours[:, k].double() - g64[:, k]).norm() / g64[:, k].norm()
    e2 = (g32[:, k].double() - g64[:, k]).norm() / g64[:, k].norm()
    print(f"{n:7s} ours-vs-o64 {e1:.2e}  o32-vs-o64 {e2:.2e}")
err = (ours.double() - g64).abs()
for idx in err.flatten().topk(8).indices:
    g, k = int(idx) // 10, int(idx) % 10
    print(f"g={g} {names[k]}: ours {ours[g,k]:.6e} o64 {g64[g,k]:.6e} o32 {g32[g,k]:.6e} radius {int(radii[0,g])} tiles {int(tt[g])} opac {opac[g]:.4f} conic {pre['conic'][g].tolist()}")
# per-duplicate comparison for the worst Gaussian: which tile is off
gw = int(err.flatten().argmax()) // 10
print("worst gaussian", gw, "dup records:\n", dup[off[gw]:off[gw] + tt[gw], :10])
print("sum |terms| estimate (conditioning): per-dup abs sum", dup[off[gw]:off[gw] + tt[gw], :10].abs().sum(0), " total", ours[gw])
