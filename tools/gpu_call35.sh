#!/usr/bin/env bash
mkdir -p gpurun_out
{
echo "# compute-sanitizer on tools/sanitize_small.py (round-2 build: integer DC chain with row replay, row-parallel ME + me_fallback_kernel,"
echo "# 192-thread intra wavefront, step-sliced copies, CUDA graphs off so that every launch is attributed)"
ICSP_GRAPHS=0 timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | tail -4
ICSP_GRAPHS=0 timeout 500 compute-sanitizer --tool racecheck python tools/sanitize_small.py 2>&1 | tail -4
ICSP_GRAPHS=0 ICSP_ME_ROWS_G=0 ICSP_INTRA_WIDE_G=0 ICSP_DC_EPS=1 timeout 500 compute-sanitizer --tool racecheck python tools/sanitize_small.py 2>&1 | tail -4
echo "# with graphs (default)"
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | tail -3
} > gpurun_out/r02_sanitizer.txt 2>&1
cat gpurun_out/r02_sanitizer.txt
python tools/fuzz_big.py 150 11 2>&1 | tail -3 | tee gpurun_out/r02_fuzz.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -n 2 gpurun_out/r02_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'parity', d['parity_check']['ok'], 'clocks', d['clocks'])
for k, v in d['baseline_configs'].items(): print(k, {a: b for a, b in v.items() if 'fps' in a or 'equal' in a})
PY
python tools/cli_bench.py 64 1 > gpurun_out/r02_cli_bench.json 2> gpurun_out/r02_cli_bench.err; cat gpurun_out/r02_cli_bench.json
