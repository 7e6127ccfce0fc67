#!/usr/bin/env bash
# ncu --set full of the transform kernels, v1 and v2, with source-level stall sampling
mkdir -p gpurun_out
ICSP_TR_V1=1 ncu --set full --clock-control none --import-source on -k regex:'fdct_quant_kernel|idct_recon_kernel' -s 4 -c 2 -o gpurun_out/tr_v1 -f python tools/prof_encode.py > gpurun_out/c3_v1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fdct_quant_kernel2|idct_recon_kernel2' -s 4 -c 2 -o gpurun_out/tr_v2 -f python tools/prof_encode.py > gpurun_out/c3_v2.log 2>&1
tail -3 gpurun_out/c3_v1.log gpurun_out/c3_v2.log
