#!/usr/bin/env bash
mkdir -p gpurun_out
N=${NGPU:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_n$N.log
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.log").read().strip().splitlines()[-1])
print("N=$N", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
