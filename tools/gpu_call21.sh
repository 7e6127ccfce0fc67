#!/usr/bin/env bash
mkdir -p gpurun_out
# one intra step + one inter step of the reduced batch (24 streams x 100 frames = 240 GOPs per launch), every kernel once, + the entropy kernels
ncu --set full --clock-control none --import-source on -s 0 -c 12 -o /tmp/r02_full -f python tools/prof_encode.py --entropy > gpurun_out/c21_full.log 2>&1
tail -n 2 gpurun_out/c21_full.log
python tools/ncu_summary.py /tmp/r02_full.ncu-rep gpurun_out/r02_ncu_full_summary.csv
python tools/ncu_stalls.py /tmp/r02_full.ncu-rep > gpurun_out/r02_ncu_stalls.txt 2>&1
ncu --set full --clock-control none -k regex:'entropy|plane_sse|parse_rows|row_index' -c 8 -o /tmp/r02_full_en -f python tools/prof_encode.py --entropy --bits-decode > gpurun_out/c21_full_en.log 2>&1
python tools/ncu_summary.py /tmp/r02_full_en.ncu-rep gpurun_out/r02_ncu_full_summary_entropy.csv
ls -la /tmp/r02_full.ncu-rep
cp /tmp/r02_full.ncu-rep gpurun_out/r02_full.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file /tmp/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/c21_bench_under_ncu.json 2> gpurun_out/c21_bench_under_ncu.err
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
for v in tmp_variants/iw_10.so tmp_variants/iw_12.so tmp_variants/iw_13.so; do cp icspcodec_b200/libicspcuda.so /tmp/keep.so; cp $v icspcodec_b200/libicspcuda.so; echo "$v: $(python tools/value_only.py 2>&1 | tail -1)"; cp /tmp/keep.so icspcodec_b200/libicspcuda.so; done
echo "default: $(python tools/value_only.py 2>&1 | tail -1)"
echo "no graphs: $(ICSP_GRAPHS=0 python tools/value_only.py 2>&1 | tail -1)"
python tools/cli_bench.py 64 1 > gpurun_out/r02_cli_bench.json 2> gpurun_out/r02_cli_bench.err; cat gpurun_out/r02_cli_bench.json; tail -n 3 gpurun_out/r02_cli_bench.err
