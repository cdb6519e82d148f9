#!/usr/bin/env bash
python -m pytest tests -m gpu -q -k "vggt" 2>&1 | grep -E "passed|failed" | tail -2
bash scripts/gpu_sanitize.sh
