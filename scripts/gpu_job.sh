#!/usr/bin/env bash
# ad-hoc GPU job: edit freely
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
bash scripts/gpu_all_workloads.sh
