#!/usr/bin/env bash
# GPU-box job of round 1: parity tests, ncu launch list + full captures, bench on every workload.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blend_ -s 6 -c 2 -o gpurun_out/prof_blend_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:project_ -s 6 -c 2 -o gpurun_out/prof_project_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
for wl in c2p c2 c3 c4; do python bench.py --steps 20 --warmup 5 --workload $wl --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_$wl.log; done
