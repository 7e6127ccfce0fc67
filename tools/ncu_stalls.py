#!/usr/bin/env python
"""print duration, occupancy, pipe utilisation and the stall-reason breakdown of every kernel in an .ncu-rep"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum']
want += [h for h in hdr if 'average_warps_issue_stalled' in h and 'not_issued' not in h]
for r in rows[2:]:
    print(r[idx['Kernel Name']][:60])
    for w in want:
        if w not in idx: continue
        v = r[idx[w]]
        try:
            if float(v) > 0.05: print('    %-52s %s' % (w.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', '')[:52], v))
        except ValueError:
            pass
