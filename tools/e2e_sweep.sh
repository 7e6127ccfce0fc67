#!/usr/bin/env bash
for cg in 240 120 80 60; do
  ICSP_CHUNK_GOPS=$cg python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('chunk $cg value %.0f e2e %.0f e2e_ms %.1f syn %.0f' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_syntax']['value']))"
done
