#!/usr/bin/env bash
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
bash scripts/gpu_all_workloads.sh
