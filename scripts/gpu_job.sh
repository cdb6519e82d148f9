#!/usr/bin/env bash
python -m pytest tests/test_raster_gpu.py -m gpu -q -x -k full_size 2>&1 | tail -25
