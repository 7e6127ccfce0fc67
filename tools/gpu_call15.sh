#!/usr/bin/env bash
for e in none chain intra chain,intra me chain,intra,me; do
  echo "skip $e: $(ICSP_EXP_SKIP=$e python tools/value_only.py 2>&1 | tail -1)"
done
