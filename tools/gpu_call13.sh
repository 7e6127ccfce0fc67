#!/usr/bin/env bash
mkdir -p gpurun_out
echo "-- kernel times"; python tools/kernel_times.py 2>&1 | tail -1
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
