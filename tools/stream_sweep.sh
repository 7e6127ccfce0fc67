#!/usr/bin/env bash
# sweep compute-stream count / chunk size: prints value per setting (resident path)
for cfg in "1 1920" "2 960" "2 480" "3 640" "4 480" "2 640"; do
  set -- $cfg
  ICSP_STREAMS=$1 ICSP_CHUNK_GOPS=$2 python tools/value_only.py
done
