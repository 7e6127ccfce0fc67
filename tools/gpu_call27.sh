#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/kernel_times.py 2>&1 | tail -1
ICSP_KT_STREAMS=1 python tools/kernel_times.py 2>&1 | tail -1
ICSP_KT_STREAMS=1 ICSP_ME_PERSISTENT=0 python tools/kernel_times.py 2>&1 | tail -1
ICSP_KT_STREAMS=4 python tools/kernel_times.py 2>&1 | tail -1
ICSP_KT_STREAMS=4 ICSP_ME_PERSISTENT=0 python tools/kernel_times.py 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline 2> gpurun_out/c27_bench.err | cut -c1-400
