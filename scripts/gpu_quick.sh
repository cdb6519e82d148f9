#!/usr/bin/env bash
# Quick GPU check: parity tests + headline bench (stage times in the JSON line).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_quick.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])
print(d['stage_ms'])
"
