#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
ICSP_INTRA_WIDE_G=1000 python tools/kernel_times.py 2>&1 | tail -1 | cut -c1-300
python tools/kernel_times.py 2>&1 | tail -1 | cut -c1-300
ICSP_KT_STREAMS=16 ICSP_INTRA_WIDE_G=1000 python tools/kernel_times.py 2>&1 | tail -1 | cut -c1-200
ICSP_KT_STREAMS=16 ICSP_INTRA_WIDE_G=0 python tools/kernel_times.py 2>&1 | tail -1 | cut -c1-200
