#!/usr/bin/env bash
python -m pytest tests/test_raster_gpu.py -m gpu -q -k "odd_sizes" 2>&1 | grep -E "passed|failed|Error|assert|rel err" | tail -6
