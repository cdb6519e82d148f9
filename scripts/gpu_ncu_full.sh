#!/usr/bin/env bash
# ncu --set full of one whole eager step of the headline workload (13 raster-path kernels) -> gpurun_out/prof_r2_final.ncu-rep
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on \
    -k regex:"blend_|project_|scan_kernel|emit_kernel|tile_sort|pose_reduce|camera_|image_mse" -s 78 -c 13 \
    -o gpurun_out/prof_r2_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-rope --no-head > gpurun_out/ncu_full.log 2>&1
grep -c "Profiling" gpurun_out/ncu_full.log; tail -2 gpurun_out/ncu_full.log | cut -c1-120
