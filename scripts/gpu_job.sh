#!/usr/bin/env bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"blend_forward|blend_backward_log|project_forward|project_backward|tile_sort_pack|emit_kernel" -s 30 -c 6 -o gpurun_out/prof_r1b -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_b.log 2>&1
tail -2 gpurun_out/ncu_full_b.log | cut -c1-200
