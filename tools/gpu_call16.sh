#!/usr/bin/env bash
for sk in 0 1 2 3 4 5 6 8 12; do
  echo "skew $sk: $(ICSP_SKEW=$sk python tools/value_only.py 2>&1 | tail -1)"
done
for cfg in "6 320 2" "8 240 2" "8 240 4" "6 320 4"; do set -- $cfg
  echo "streams $1 chunk $2 skew $3: $(ICSP_STREAMS=$1 ICSP_CHUNK_GOPS=$2 ICSP_SKEW=$3 python tools/value_only.py 2>&1 | tail -1)"
done
