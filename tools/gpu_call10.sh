#!/usr/bin/env bash
for cfg in "22 0 4" "11 0 4" "11 118000 4" "11 118000 6" "11 118000 8" "8 118000 6" "6 118000 6"; do
  set -- $cfg
  echo "seg $1 smem_min $2 streams $3: $(ICSP_ME_SEG=$1 ICSP_ME_SMEM_MIN=$2 ICSP_STREAMS=$3 python tools/value_only.py 2>&1 | tail -1)"
done
