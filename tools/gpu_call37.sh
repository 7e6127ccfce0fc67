#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for s in 64 1 8; do
  echo "streams $s fused:   $(ICSP_KT_STREAMS=$s python tools/kernel_times.py 2>&1 | tail -1 | cut -c1-110)"
  echo "streams $s unfused: $(ICSP_KT_STREAMS=$s ICSP_FUSE_CHAIN=0 python tools/kernel_times.py 2>&1 | tail -1 | cut -c1-110)"
done
