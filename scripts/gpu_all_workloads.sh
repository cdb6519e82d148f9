#!/usr/bin/env bash
# bench.py on every workload (no CPU baseline) -> gpurun_out/bench_<wl>.log
mkdir -p gpurun_out
for wl in ${WLS:-c2p c2 c3 c4}; do
  python bench.py --steps 20 --warmup 6 --workload $wl --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$wl.log
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$wl.log").read())
print("$wl", "views/s", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "N", d["config"]["duplicates_per_step"])
print("   ", d["stage_ms"], d.get("pair_log"))
PY
done
