#!/usr/bin/env bash
# The driver's scaling series on ONE box: bench.py at N = 1, 2, 4, 8 back to back -> gpurun_out/series_n*.json
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-rope --no-head 2>/dev/null | tail -1 > gpurun_out/series_n1.json
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n \
      bench.py --gpus $n --steps 20 --warmup 5 2>gpurun_out/series_n$n.err | tail -1 > gpurun_out/series_n$n.json
done
python - <<PY
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/series_n{n}.json").read())
    except Exception as exc:
        print(n, "FAILED", exc); continue
    base = base or d["value"]
    e = d["e2e"]
    print(f"N={n} value {d['value']:.0f} ({d['ms_per_step']} ms) eff {d['value'] / (base * n):.3f}  e2e {e['value']:.0f}  copy GB/s/rank {e['h2d_copy_only_gbs_per_rank']}")
PY
