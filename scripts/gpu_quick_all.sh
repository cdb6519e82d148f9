#!/usr/bin/env bash
bash scripts/gpu_quick.sh
bash scripts/gpu_all_workloads.sh
