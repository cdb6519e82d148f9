#!/usr/bin/env bash
python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_full.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k: d[k] for k in ('value','ms_per_step','e2e','roofline','cpu_baseline','clocks','gpu_launches')})
print(d['stage_ms'])"
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-400
