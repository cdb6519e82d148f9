#!/usr/bin/env bash
# Round-end evidence pass (1 GPU): GPU parity tests, smoke, default bench, reference arm, all workloads, ncu launch list,
# ncu --set full of the raster-path kernels (headline step) and of the raw-head projection kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed" | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.log | cut -c1-200
python bench.py --impl reference --steps 4 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference.log; cut -c1-160 gpurun_out/bench_reference.log
for w in c2 c3 c4; do python bench.py --workload $w --steps 10 --warmup 4 --no-cpu-baseline --no-rope --no-head 2>/dev/null | tail -1 > gpurun_out/bench_$w.log; cut -c1-140 gpurun_out/bench_$w.log; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-rope --no-head > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"spf::" -s 78 -c 13 -o gpurun_out/prof_r2_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-rope --no-head > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-120
ncu --set full --clock-control none --import-source on -k regex:"raw_kernel" -s 8 -c 2 -o gpurun_out/prof_r2_head -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-rope > gpurun_out/ncu_head.log 2>&1
tail -2 gpurun_out/ncu_head.log | cut -c1-120
