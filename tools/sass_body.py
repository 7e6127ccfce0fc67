#!/usr/bin/env python
"""print the SASS of the per-block loop of a kernel (see sass_loop.py); usage: sass_body.py <kernel-substr> [nofp64]"""
import re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "icspcodec_b200", "libicspcuda.so")
txt = subprocess.check_output(["cuobjdump", "-sass", lib], text=True)
on = False; ins = []
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if on and ins: break
        on = sys.argv[1] in m.group(1)
        continue
    if not on: continue
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", line)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
loop = None
for pc, t in ins:
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < pc:
        tgt = int(m.group(1), 16)
        if loop is None or pc - tgt > loop[1] - loop[0]: loop = (tgt, pc)
for pc, t in ins:
    if loop[0] <= pc <= loop[1]:
        if len(sys.argv) > 2 and re.match(r"(@!?U?P\d+\s+)?D(ADD|MUL|FMA)", t): continue
        print("%04x  %s" % (pc, t))
