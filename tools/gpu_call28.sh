#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/kernel_times.py 2>&1 | tail -1
ICSP_KT_STREAMS=1 python tools/kernel_times.py 2>&1 | tail -1
for s in 2 3; do for r in 0 1000; do echo "streams $s rows_g $r"; ICSP_KT_STREAMS=$s ICSP_ME_ROWS_G=$r python tools/kernel_times.py 2>&1 | tail -1 | cut -c1-120; done; done
