#!/usr/bin/env bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -n 2 gpurun_out/r02_bench.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-300 gpurun_out/r02_bench_reference.json
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
