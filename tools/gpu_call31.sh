#!/usr/bin/env bash
# chunk-size sweep for mid-size batches (resident value; 4 compute streams)
for s in 2 4 8 16 32; do
  for cg in 0 30 40 60 80 120 160; do
    if [ "$cg" = "0" ]; then r=$(ICSP_KT_STREAMS=$s python tools/kernel_times.py 2>&1 | tail -1 | cut -c1-45);
    else r=$(ICSP_KT_STREAMS=$s ICSP_CHUNK_GOPS=$cg python tools/kernel_times.py 2>&1 | tail -1 | cut -c1-45); fi
    echo "streams $s chunk $cg: $r"
  done
done
