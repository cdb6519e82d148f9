#!/usr/bin/env bash
WLS="c2p" bash scripts/gpu_all_workloads.sh
