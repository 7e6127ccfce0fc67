#!/usr/bin/env bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'fdct_quant_kernel2|idct_recon_kernel2' -s 4 -c 2 -o gpurun_out/tr_v2b -f python tools/prof_encode.py > gpurun_out/c5_v2.log 2>&1
tail -n 3 gpurun_out/c5_v2.log
