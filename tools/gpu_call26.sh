#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/kernel_times.py 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline 2> gpurun_out/c26_bench.err | cut -c1-1500
ICSP_GRAPHS=0 timeout 420 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | tail -4
ICSP_GRAPHS=0 timeout 420 compute-sanitizer --tool racecheck python tools/sanitize_small.py 2>&1 | tail -4
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | tail -3
