#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "full_size or baseline_configs" 2>&1 | tail -5
( time python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err ) 2>&1 | tail -4
tail -n 5 gpurun_out/r02_bench_a.err
