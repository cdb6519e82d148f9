"""torchrun script: correctness and stand-alone timing of the in-switch (NVLS multimem) all-reduce against NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/nvls_check.py [--mib 64] [--out gpurun_out/nvls_check.json]

Rank 0 prints one JSON line: {"world", "nvls": bool, "error", "max_abs_diff_vs_nccl", "us": {"nccl": .., "nvls_b8": ..}}.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spfsplatv2_b200.dp import GradAllReduce  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=64)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    numel = args.mib * 1024 * 1024 // 4
    red = GradAllReduce(dev)
    buf = red.alloc(numel)
    res = {"world": world, "mib": args.mib, "nvls": red.uses_nvls(buf), "error": red.nvls_error, "us": {}}

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    src = torch.randn(numel, device=dev, generator=g)
    ref = src.clone()
    dist.all_reduce(ref)

    def time_it(r, fn, iters=20):
        for _ in range(3):
            fn()
        r.wait()
        dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        r.wait()
        e1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    plain = torch.zeros(numel, device=dev)
    nccl = GradAllReduce(dev, backend="nccl")
    res["us"]["nccl"] = round(time_it(nccl, lambda: nccl.launch([plain])), 1)
    if res["nvls"]:
        buf.copy_(src)
        red.launch([buf])
        red.wait()
        torch.cuda.synchronize(dev)
        diff = (buf - ref).abs().max()
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        res["max_abs_diff_vs_nccl"] = float(diff)
        # all ranks must hold bit-identical sums
        chk = buf.double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        res["ranks_identical"] = bool(lo.item() == hi.item())
        buf.zero_()
        for nb in (4, 8, 16, 32):
            red.nvls_blocks = nb
            res["us"][f"nvls_b{nb}"] = round(time_it(red, lambda: red.launch([buf])), 1)
    if rank == 0:
        print(json.dumps(res))
        if args.out:
            os.makedirs(os.path.dirname(args.out), exist_ok=True)
            with open(args.out, "w") as f:
                json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
