#!/usr/bin/env bash
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/kernel_times.py 2>&1 | tail -1 | cut -c1-420
ICSP_KT_STREAMS=1 python tools/kernel_times.py 2>&1 | tail -1 | cut -c1-300
