#!/bin/bash
# usage: tools/variant_sweep.sh tmp_variants/a.so tmp_variants/b.so ...   (GPU box; compares `value` of prebuilt libraries)
cp icspcodec_b200/libicspcuda.so /tmp/keep.so
for v in "$@"; do
  cp "$v" icspcodec_b200/libicspcuda.so
  echo "== $v"; python tools/value_only.py 2>&1 | tail -1
done
cp /tmp/keep.so icspcodec_b200/libicspcuda.so
