#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -5
ncu --set full --clock-control none --import-source on -k regex:'dc_chain_kernel' -s 3 -c 2 -o gpurun_out/chain_v2 -f python tools/prof_encode.py > gpurun_out/c14_ncu.log 2>&1
tail -n 1 gpurun_out/c14_ncu.log
