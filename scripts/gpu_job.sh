#!/usr/bin/env bash
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
WLS="c2p c2" bash scripts/gpu_all_workloads.sh
