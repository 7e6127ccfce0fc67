#!/usr/bin/env bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; tail -n 2 gpurun_out/r02_bench_2gpu.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_2gpu.json').read().strip().splitlines()[-1])
print('N=2 value', d['value'], 'e2e', d['e2e']['value'], 'strong', d.get('strong_scaling'), 'parity', d.get('parity_check', {}).get('ok'))
for k, v in (d.get('baseline_configs') or {}).items(): print(k, {a: b for a, b in v.items() if 'fps' in a or 'equal' in a})
PY
