#!/usr/bin/env python
"""SASS opcode histogram per kernel of libicspcuda.so (cuobjdump -sass), written as profiles/<tag>_sass_hist.json / .md.

    python tools/sass_hist.py [tag] [kernel-substring ...]

Static counts (instructions in the binary, loops counted once).  For the transform kernels the per-block loop body is
the whole kernel apart from a short prologue, so "non-FP64 per FP64" of the static listing is the ratio the warp
scheduler sees per 8x8 block (the `#pragma unroll 1` loop over the 6 blocks of a macroblock executes the body once
per block)."""
from __future__ import annotations

import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "icspcodec_b200", "libicspcuda.so")
FP64 = ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX")
CONV64 = ("I2F.F64", "F2I.F64", "F2I.S32.F64", "F2I.U32.F64", "F2F.F64", "I2F.F64.S32", "F2I.FLOOR", "F2I.TRUNC")


def demangle(name: str) -> str:
    try:
        return subprocess.check_output(["c++filt", name], text=True).strip().split("(")[0]
    except Exception:
        return name


def parse(lib: str):
    txt = subprocess.check_output(["cuobjdump", "-sass", lib], text=True)
    kernels: dict[str, list[str]] = {}
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = demangle(m.group(1))
            kernels[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            kernels[cur].append(m.group(1))
    return kernels


def summarize(ops: list[str]) -> dict:
    h = collections.Counter(ops)
    base = collections.Counter(o.split(".")[0] for o in ops)
    fp64 = sum(v for k, v in base.items() if k in FP64)
    conv = sum(v for k, v in h.items() if (k.startswith("I2F") or k.startswith("F2I") or k.startswith("F2F")) and "64" in k)
    total = len(ops)
    return {"total": total, "fp64": fp64, "conv64": conv, "non_fp64": total - fp64,
            "non_fp64_per_fp64": round((total - fp64) / fp64, 3) if fp64 else None,
            "by_base": dict(base.most_common()), "by_full": dict(h.most_common(60))}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    pats = sys.argv[2:]
    ks = parse(LIB)
    out = {}
    for name, ops in ks.items():
        if pats and not any(p in name for p in pats):
            continue
        out[name] = summarize(ops)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", f"{tag}_sass_hist.json"), "w") as f:
        json.dump(out, f, indent=1)
    for name, s in sorted(out.items(), key=lambda kv: -kv[1]["total"]):
        top = ", ".join(f"{k} {v}" for k, v in list(s["by_base"].items())[:14])
        print(f"{name}: total {s['total']}, FP64 {s['fp64']}, conv64 {s['conv64']}, non-FP64/FP64 {s['non_fp64_per_fp64']}\n    {top}")


if __name__ == "__main__":
    main()
