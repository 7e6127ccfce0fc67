#!/usr/bin/env python
"""top stall-sample instructions of one kernel in an .ncu-rep: ncu_hot.py rep kernel-regex [n]"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2]], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in body)
execd = sum(int(r[ix["Instructions Executed"]]) for r in body)
print("instructions:", len(body), "samples:", tot, "warp-instr executed:", execd)
stall_cols = [h for h in hdr if h.startswith("stall_")]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]]))[:n]
for i in sorted(order):
    r = body[i]
    top = sorted(((int(r[ix[c]]), c) for c in stall_cols if r[ix[c]].isdigit()), reverse=True)[:2]
    print("%5d %5.1f%%  %-60s %s" % (i, 100.0 * int(r[ix["# Samples"]]) / tot, r[ix["Source"]].strip()[:60], " ".join(f"{c[6:]}={v}" for v, c in top if v)))
