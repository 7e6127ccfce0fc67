"""Builds libicspcuda.so (and the host tools) IN-TREE for sm_100a with nvcc; no JIT cache is used."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
LIB = os.path.join(PKG, "libicspcuda.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
              "-fmad=false",           # never contract a*b+c: the reference rounds every product
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libicspcuda cannot be built (there is no CPU fallback)")


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force: bool = False, verbose: bool = False) -> str:
    """Idempotent and safe under several ranks starting at once (file lock)."""
    import fcntl
    with open(os.path.join(PKG, ".build.lock"), "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> str:
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, "include", "icspcuda.h")]
    if force or _stale(LIB, srcs):
        cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-o", LIB, os.path.join(CSRC, "icspcuda.cu")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    host_srcs = [os.path.join(HOST, f) for f in sorted(os.listdir(HOST)) if f.endswith((".cpp", ".h"))] if os.path.isdir(HOST) else []
    if host_srcs and os.path.exists(os.path.join(HOST, "Makefile")):
        subprocess.check_call(["make", "-s", "-C", HOST] + (["-B"] if force else []))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
