#!/usr/bin/env bash
for mc in 8 16 32; do
  echo "maxconn $mc: $(CUDA_DEVICE_MAX_CONNECTIONS=$mc python tools/value_only.py 2>&1 | tail -1)"
done
echo "maxconn 32 hi_prio 0: $(CUDA_DEVICE_MAX_CONNECTIONS=32 ICSP_HI_PRIO=0 python tools/value_only.py 2>&1 | tail -1)"
for cfg in "6 320" "8 240" "3 640" "2 960"; do set -- $cfg
  echo "maxconn 32 streams $1 chunk $2: $(CUDA_DEVICE_MAX_CONNECTIONS=32 ICSP_STREAMS=$1 ICSP_CHUNK_GOPS=$2 python tools/value_only.py 2>&1 | tail -1)"
done
echo "maxconn 32 skew 2: $(CUDA_DEVICE_MAX_CONNECTIONS=32 ICSP_SKEW=2 python tools/value_only.py 2>&1 | tail -1)"
