#!/usr/bin/env bash
mkdir -p gpurun_out
echo "-- v2 two-loop"; python tools/kernel_times.py 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:'fdct_quant_kernel2|idct_recon_kernel2' -s 4 -c 2 -o gpurun_out/tr_v2d -f python tools/prof_encode.py > gpurun_out/c7_v2.log 2>&1
tail -n 1 gpurun_out/c7_v2.log
