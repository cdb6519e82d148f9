#!/usr/bin/env bash
python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|Error|assert" | tail -6
