#!/usr/bin/env bash
python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|Error|error" | tail -5
bash scripts/gpu_all_workloads.sh
