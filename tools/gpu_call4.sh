#!/usr/bin/env bash
mkdir -p gpurun_out
cp icspcodec_b200/libicspcuda.so /tmp/keep.so
echo "-- v1 (ICSP_TR_V1=1)"; ICSP_TR_V1=1 python tools/kernel_times.py 2>&1 | tail -1
for v in tmp_variants/v2_*.so; do
  cp "$v" icspcodec_b200/libicspcuda.so
  echo "-- $v"; python tools/kernel_times.py 2>&1 | tail -1
done
cp /tmp/keep.so icspcodec_b200/libicspcuda.so
