#!/usr/bin/env python
"""Static opcode histogram of the innermost per-block loop of a kernel (cuobjdump -sass): the region between the target of
the last backward branch and that branch.  usage: sass_loop.py <kernel-substring> [lib]"""
import collections, re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "icspcodec_b200", "libicspcuda.so")
txt = subprocess.check_output(["cuobjdump", "-sass", lib], text=True)
on = False; ins = []
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        on = sys.argv[1] in m.group(1)
        if on and ins: break
        continue
    if not on: continue
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", line)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
# last backward branch
loop = None
for i, (pc, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < pc and "EXIT" not in t:
        tgt = int(m.group(1), 16)
        if loop is None or pc - tgt > loop[1] - loop[0]: loop = (tgt, pc)
print("kernel instrs:", len(ins), "loop:", [hex(x) for x in loop] if loop else None)
body = [t for pc, t in ins if loop and loop[0] <= pc <= loop[1]]
def base(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    return t.split()[0].split(".")[0]
h = collections.Counter(base(t) for t in body)
fp64 = sum(h[k] for k in ("DADD", "DMUL", "DFMA", "DSETP"))
print("loop instrs:", len(body), "FP64:", fp64, "non-FP64:", len(body) - fp64, "ratio: %.2f" % ((len(body) - fp64) / max(fp64, 1)))
print(", ".join(f"{k} {v}" for k, v in h.most_common()))
