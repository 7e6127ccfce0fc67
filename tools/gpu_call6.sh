#!/usr/bin/env bash
mkdir -p gpurun_out
echo "-- v1"; ICSP_TR_V1=1 python tools/kernel_times.py 2>&1 | tail -1
echo "-- v2 rolled"; python tools/kernel_times.py 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:'fdct_quant_kernel2|idct_recon_kernel2' -s 4 -c 2 -o gpurun_out/tr_v2c -f python tools/prof_encode.py > gpurun_out/c6_v2.log 2>&1
tail -n 2 gpurun_out/c6_v2.log
