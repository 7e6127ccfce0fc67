#!/usr/bin/env bash
mkdir -p gpurun_out
echo "-- v2 two-loop + deferred shifts + border-first"; python tools/kernel_times.py 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
