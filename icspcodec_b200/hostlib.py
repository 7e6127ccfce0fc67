"""ctypes binding of icspcodec_b200/host/libicsphost.so — the C++ bitstream writer/reader used by icspenc/icspdec."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def load() -> C.CDLL:
    global _LIB
    if _LIB is None:
        p = os.path.join(PKG, "host", "libicsphost.so")
        if not os.path.exists(p):
            raise RuntimeError(f"{p} missing: run `python -m icspcodec_b200.build`")
        _LIB = C.CDLL(p)
        _LIB.icsp_host_write_stream.restype = C.c_long
        _LIB.icsp_host_write_stream_indexed.restype = C.c_long
    return _LIB


def _p(a):
    return None if a is None else np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)


def write_stream(levels, acflag, mpm, ipm, mvd, w: int, h: int, qdc: int, qac: int, ip: int, threads: int = 4) -> bytes:
    n = levels.shape[0]
    cap = 64 + n * w * h * 4
    out = np.zeros(cap, np.uint8)
    arrs = [np.ascontiguousarray(levels, np.int16), np.ascontiguousarray(acflag, np.uint8), np.ascontiguousarray(mpm, np.uint8),
            np.ascontiguousarray(ipm, np.uint8), np.ascontiguousarray(mvd, np.int16)]
    ln = load().icsp_host_write_stream(*[_p(a) for a in arrs], n, w, h, qdc, qac, ip, threads, _p(out), C.c_long(cap))
    if ln < 0:
        raise ValueError("icsp_host_write_stream failed")
    return out[:ln].tobytes()


def parse_stream(data: bytes, nframes: int):
    arr = np.frombuffer(data, np.uint8)
    hdr = (C.c_int * 5)()
    if load().icsp_host_parse_stream(_p(arr), C.c_long(len(data)), 0, hdr, None, None, None, None, None) != 0:
        raise ValueError("bad stream header")
    w, h, qdc, qac, ip = list(hdr)
    nmb = (w // 16) * (h // 16)
    levels = np.zeros((nframes, nmb, 6, 64), np.int16)
    acflag = np.zeros((nframes, nmb, 6), np.uint8)
    mpm = np.zeros((nframes, nmb, 4), np.uint8)
    ipm = np.zeros((nframes, nmb, 4), np.uint8)
    mvd = np.zeros((nframes, nmb, 2), np.int16)
    rc = load().icsp_host_parse_stream(_p(arr), C.c_long(len(data)), nframes, hdr, _p(levels), _p(acflag), _p(mpm), _p(ipm), _p(mvd))
    if rc != 0:
        raise ValueError("stream parse failed")
    return dict(levels=levels, acflag=acflag, mpm=mpm, ipm=ipm, mvd=mvd), dict(w=w, h=h, qdc=qdc, qac=qac, ip=ip)


def write_stream_indexed(levels, acflag, mpm, ipm, mvd, w: int, h: int, qdc: int, qac: int, ip: int, threads: int = 4):
    """-> (file bytes, rows[n][h/16] uint64): the stream plus the bit offset of every macroblock row from the body start."""
    n = levels.shape[0]
    cap = 64 + n * (w // 16) * (h // 16) * 800
    out = np.zeros(cap, np.uint8)
    rows = np.zeros((n, h // 16), np.uint64)
    arrs = [np.ascontiguousarray(levels, np.int16), np.ascontiguousarray(acflag, np.uint8), np.ascontiguousarray(mpm, np.uint8),
            np.ascontiguousarray(ipm, np.uint8), np.ascontiguousarray(mvd, np.int16)]
    ln = load().icsp_host_write_stream_indexed(*[_p(a) for a in arrs], n, w, h, qdc, qac, ip, threads, _p(out), C.c_long(cap), _p(rows))
    if ln < 0:
        raise ValueError("icsp_host_write_stream_indexed failed")
    return out[:ln].tobytes(), rows


def parse_stream_indexed(data: bytes, nframes: int, rows: np.ndarray):
    """Parse with every macroblock row restarted at its recorded bit offset (the chains of the GPU bit reader)."""
    arr = np.frombuffer(data, np.uint8)
    hdr = (C.c_int * 5)()
    if load().icsp_host_parse_stream(_p(arr), C.c_long(len(data)), 0, hdr, None, None, None, None, None) != 0:
        raise ValueError("bad stream header")
    w, h, qdc, qac, ip = list(hdr)
    nmb = (w // 16) * (h // 16)
    levels = np.zeros((nframes, nmb, 6, 64), np.int16)
    acflag = np.zeros((nframes, nmb, 6), np.uint8)
    mpm = np.zeros((nframes, nmb, 4), np.uint8)
    ipm = np.zeros((nframes, nmb, 4), np.uint8)
    mvd = np.zeros((nframes, nmb, 2), np.int16)
    rows = np.ascontiguousarray(rows, np.uint64)
    rc = load().icsp_host_parse_stream_indexed(_p(arr), C.c_long(len(data)), nframes, _p(rows), C.c_long(rows.size), _p(levels), _p(acflag),
                                               _p(mpm), _p(ipm), _p(mvd))
    if rc != 0:
        raise ValueError("indexed stream parse failed")
    return dict(levels=levels, acflag=acflag, mpm=mpm, ipm=ipm, mvd=mvd), dict(w=w, h=h, qdc=qdc, qac=qac, ip=ip)
