#!/usr/bin/env bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'idct_recon_kernel2|me_sad_frame' -s 6 -c 2 -o gpurun_out/tr_v2e -f python tools/prof_encode.py > gpurun_out/c9.log 2>&1
tail -n 1 gpurun_out/c9.log
