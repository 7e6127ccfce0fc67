#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c33_bench.json 2> gpurun_out/c33_bench.err; tail -n 2 gpurun_out/c33_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/c33_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'parity', d['parity_check']['ok'])
for k, v in d['baseline_configs'].items(): print(k, {a: b for a, b in v.items() if 'fps' in a or 'equal' in a})
PY
ICSP_SLICE_COPIES=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c33_bench_noslice.json 2> gpurun_out/c33_bench_noslice.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/c33_bench_noslice.json').read().strip().splitlines()[-1])
for k, v in d['baseline_configs'].items(): print('noslice', k, {a: b for a, b in v.items() if 'fps' in a})
PY
