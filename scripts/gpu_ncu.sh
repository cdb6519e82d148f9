#!/usr/bin/env bash
# ncu --set full capture of kernels matching $1 (regex) during a short headline bench; report -> gpurun_out/prof_$2.ncu-rep
# usage: bash scripts/gpu_ncu.sh <kernel-regex> <tag> [skip] [count] [workload]
mkdir -p gpurun_out
REGEX=$1; TAG=$2; SKIP=${3:-3}; COUNT=${4:-1}; WL=${5:-c2p}
ncu --set full --clock-control none --import-source on -k regex:$REGEX -s $SKIP -c $COUNT -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $WL > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log | cut -c1-300
