"""Micro-benchmark of the sm_100a RoPE kernel on SPFSplatV2's shapes: achieved GB/s vs measured HBM peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spfsplatv2_b200.curope import rope_2d

peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
d = "cuda:0"
out = []
for name, (B, N, H, D), dt in [("encoder q/k (48,256,16,64) fp32", (48, 256, 16, 64), torch.float32),
                               ("decoder q (16,258,12,64) fp32", (16, 258, 12, 64), torch.float32),
                               ("cross-attn k (16,516,12,64) fp32", (16, 516, 12, 64), torch.float32),
                               ("encoder q/k bf16", (48, 256, 16, 64), torch.bfloat16),
                               ("large (64,1024,16,64) fp32 (> L2)", (64, 1024, 16, 64), torch.float32)]:
    # tokens as the attention blocks hand them over: a [B,N,H,D] view of a fused qkv tensor [B,N,3,H,D]
    qkv = torch.randn(B, N, 3, H, D, device=d, dtype=dt)
    tok = qkv[:, :, 0]
    pos = torch.randint(0, 16, (B, N, 2), device=d)
    for _ in range(5):
        rope_2d(tok, pos, 100.0, 1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 50
    e0.record()
    for _ in range(iters):
        rope_2d(tok, pos, 100.0, 1.0)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    nbytes = 2 * B * N * H * D * tok.element_size() + 16 * B * N
    out.append(dict(case=name, us=round(us, 2), gbs=round(nbytes / us / 1e3, 1), frac_of_measured_hbm=round(nbytes / us / 1e3 / peak, 3)))
    print(out[-1])
json.dump(out, open("gpurun_out/rope_bench.json", "w"), indent=1)
