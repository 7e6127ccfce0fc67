#!/usr/bin/env bash
# round-2 GPU call 2: parity of the v2 transform kernels, A/B of the variants, fixed micro-benchmark
mkdir -p gpurun_out
tools/microbench_peaks > gpurun_out/r02_peaks.json 2> gpurun_out/c2_peaks.err
grep -E "LDS|DADD\+0|I2F|F2I" gpurun_out/r02_peaks.json
echo "== parity (gpu tests)"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== variants"
cp icspcodec_b200/libicspcuda.so /tmp/keep.so
echo "-- v1 (ICSP_TR_V1=1)"; ICSP_TR_V1=1 python tools/kernel_times.py 2>&1 | tail -1
for v in tmp_variants/v2_default.so tmp_variants/v2_MAGIC_CVT0.so tmp_variants/v2_MAGIC_FLOOR0.so tmp_variants/v2_TR_UNROLL0.so; do
  cp "$v" icspcodec_b200/libicspcuda.so
  echo "-- $v"; python tools/kernel_times.py 2>&1 | tail -1
done
cp /tmp/keep.so icspcodec_b200/libicspcuda.so
