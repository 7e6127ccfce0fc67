#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -n 2 gpurun_out/r02_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'bits-only', d['e2e_bitstream_only']['value'], 'parity', d['parity_check']['ok'], 'clocks', d['clocks'])
for k, v in d['baseline_configs'].items(): print(k, {a: b for a, b in v.items() if 'fps' in a or 'equal' in a})
PY
python tools/pcie_probe.py 2>&1 | tail -6
