#!/usr/bin/env bash
# Round-end verification pass: GPU parity tests, smoke, default bench (graph), eager bench, reference arm, all workloads.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed" | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.log | cut -c1-160
python bench.py --impl reference --steps 4 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference.log; cut -c1-120 gpurun_out/bench_reference.log
bash scripts/gpu_all_workloads.sh
