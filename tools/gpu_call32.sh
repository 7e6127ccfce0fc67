#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for s in 1 2 3 4 8 16; do echo "streams $s: $(ICSP_KT_STREAMS=$s python tools/kernel_times.py 2>&1 | tail -1 | cut -c1-45)"; done
# one intra step + one inter step of the reduced batch (24 streams x 100 frames = 240 GOPs per launch), every kernel once, + the entropy kernels
ICSP_CHUNK_GOPS=240 ICSP_INTRA_WIDE_G=0 ncu --set full --clock-control none --import-source on -s 0 -c 12 -o /tmp/r02_full -f python tools/prof_encode.py --entropy > gpurun_out/c32_full.log 2>&1
tail -n 2 gpurun_out/c32_full.log
python tools/ncu_summary.py /tmp/r02_full.ncu-rep gpurun_out/r02_ncu_full_summary.csv
python tools/ncu_stalls.py /tmp/r02_full.ncu-rep > gpurun_out/r02_ncu_stalls.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file /tmp/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/c32_bench_under_ncu.json 2> gpurun_out/c32_bench_under_ncu.err
python - <<'PY'
import csv
rows = list(csv.reader(open('/tmp/r02_launches_bench.csv', errors='ignore')))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
h = rows[hi]; ix = {k: i for i, k in enumerate(h)}
with open('gpurun_out/r02_launches_bench.csv', 'w') as f:
    f.write('id,kernel,grid,block,us\n')
    for r in rows[hi + 2:]:
        if len(r) < len(h): continue
        f.write('%s,%s,"%s","%s",%s\n' % (r[ix['ID']], r[ix['Kernel Name']].split('(')[0], r[ix['Grid Size']], r[ix['Block Size']], r[ix['Metric Value']]))
PY
wc -l gpurun_out/r02_launches_bench.csv
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -n 3 gpurun_out/r02_bench.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-300 gpurun_out/r02_bench_reference.json
python tools/cli_bench.py 64 1 > gpurun_out/r02_cli_bench.json 2> gpurun_out/r02_cli_bench.err; cat gpurun_out/r02_cli_bench.json; tail -n 3 gpurun_out/r02_cli_bench.err
