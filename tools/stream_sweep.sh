#!/usr/bin/env bash
# sweep compute-stream count / chunk size: prints value, e2e per setting
for cfg in "1 1920" "2 240" "2 480" "3 240" "4 240" "2 120"; do
  set -- $cfg
  ICSP_STREAMS=$1 ICSP_CHUNK_GOPS=$2 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('streams $1 chunk $2 value %.0f e2e %.0f ms/step %.2f e2e_ms %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']))"
done
