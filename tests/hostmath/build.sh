#!/usr/bin/env bash
# Host (g++) build of the shared per-Gaussian math for CPU unit tests.  Test infrastructure only.
set -euo pipefail
cd "$(dirname "$0")"
g++ -O2 -ffp-contract=off -shared -fPIC -o libhostmath.so hostmath.cpp
