"""TEST INFRASTRUCTURE — ctypes binding of oracle/libicsp_oracle.so plus helpers that run the
compiled, unmodified reference binaries in oracle/_ref/.  Importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; the product package
(icspcodec_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import tempfile
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_LIB = None


def build() -> None:
    subprocess.check_call(["make", "-s", "-C", HERE, "libicsp_oracle.so"])
    if os.path.isdir("/root/reference/source/encoder"):
        subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "libicsp_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.icsp_oracle_write_bitstream.restype = C.c_long
    return _LIB


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class Syntax:
    """SoA syntax + reconstruction of a run of frames (layout in icsp_oracle.h / include/icspcuda.h)."""
    levels: np.ndarray   # int16 [n][nmb][6][64]
    acflag: np.ndarray   # uint8 [n][nmb][6]
    mpm: np.ndarray      # uint8 [n][nmb][4]
    ipm: np.ndarray      # uint8 [n][nmb][4]
    mvd: np.ndarray      # int16 [n][nmb][2]
    mv: np.ndarray | None = None       # int16 [n][nmb][2]
    minsad: np.ndarray | None = None   # int32 [n][nmb]
    recon: np.ndarray | None = None    # uint8 [n][fb]
    dct: np.ndarray | None = None      # f64 [n][nmb][6][64]


def alloc_syntax(n: int, w: int, h: int, with_dct: bool = False) -> Syntax:
    nmb = (w // 16) * (h // 16)
    return Syntax(
        levels=np.zeros((n, nmb, 6, 64), np.int16), acflag=np.zeros((n, nmb, 6), np.uint8),
        mpm=np.zeros((n, nmb, 4), np.uint8), ipm=np.zeros((n, nmb, 4), np.uint8),
        mvd=np.zeros((n, nmb, 2), np.int16), mv=np.zeros((n, nmb, 2), np.int16),
        minsad=np.zeros((n, nmb), np.int32), recon=np.zeros((n, w * h * 3 // 2), np.uint8),
        dct=np.zeros((n, nmb, 6, 64), np.float64) if with_dct else None)


def encode(frames: np.ndarray, w: int, h: int, qdc: int, qac: int, ip: int, with_dct: bool = False) -> Syntax:
    frames = np.ascontiguousarray(frames, np.uint8)
    n = frames.shape[0]
    s = alloc_syntax(n, w, h, with_dct)
    rc = lib().icsp_oracle_encode(_p(frames), n, w, h, qdc, qac, ip, _p(s.levels), _p(s.acflag), _p(s.mpm), _p(s.ipm),
                                  _p(s.mvd), _p(s.mv), _p(s.minsad), _p(s.recon), _p(s.dct))
    if rc != 0:
        raise ValueError("icsp_oracle_encode failed")
    return s


def decode(s: Syntax, w: int, h: int, qdc: int, qac: int, ip: int) -> np.ndarray:
    n = s.levels.shape[0]
    out = np.zeros((n, w * h * 3 // 2), np.uint8)
    rc = lib().icsp_oracle_decode(_p(s.levels), _p(s.mpm), _p(s.ipm), _p(s.mvd), n, w, h, qdc, qac, ip, _p(out))
    if rc != 0:
        raise ValueError("icsp_oracle_decode failed")
    return out


def write_bitstream(s: Syntax, w: int, h: int, qdc: int, qac: int, ip: int) -> bytes:
    n = s.levels.shape[0]
    cap = 64 + n * w * h * 4
    buf = np.zeros(cap, np.uint8)
    ln = lib().icsp_oracle_write_bitstream(_p(s.levels), _p(s.acflag), _p(s.mpm), _p(s.ipm), _p(s.mvd), n, w, h, qdc,
                                           qac, ip, _p(buf), C.c_long(cap))
    if ln < 0:
        raise ValueError("icsp_oracle_write_bitstream failed")
    return buf[:ln].tobytes()


def parse_bitstream(data: bytes, nframes: int):
    arr = np.frombuffer(data, np.uint8)
    w, h, qdc, qac, ip = (C.c_int() for _ in range(5))
    rc = lib().icsp_oracle_parse_bitstream(_p(arr), C.c_long(len(data)), nframes, C.byref(w), C.byref(h), C.byref(qdc),
                                           C.byref(qac), C.byref(ip), None, None, None, None, None)
    if rc != 0:
        raise ValueError("bad header")
    s = alloc_syntax(nframes, w.value, h.value)
    rc = lib().icsp_oracle_parse_bitstream(_p(arr), C.c_long(len(data)), nframes, C.byref(w), C.byref(h), C.byref(qdc),
                                           C.byref(qac), C.byref(ip), _p(s.levels), _p(s.acflag), _p(s.mpm), _p(s.ipm),
                                           _p(s.mvd))
    if rc != 0:
        raise ValueError("parse failed")
    s.mv = s.minsad = s.recon = None
    return s, dict(w=w.value, h=h.value, qdc=qdc.value, qac=qac.value, ip=ip.value)


def dct8x8(blocks: np.ndarray) -> np.ndarray:
    blocks = np.ascontiguousarray(blocks, np.int32).reshape(-1, 64)
    out = np.zeros(blocks.shape, np.float64)
    lib().icsp_oracle_dct8x8(_p(blocks), _p(out), blocks.shape[0])
    return out


def idct8x8(blocks: np.ndarray, table: int = 0) -> np.ndarray:
    blocks = np.ascontiguousarray(blocks, np.int32).reshape(-1, 64)
    out = np.zeros(blocks.shape, np.float64)
    lib().icsp_oracle_idct8x8(_p(blocks), _p(out), blocks.shape[0], table)
    return out


def me(cur_y: np.ndarray, ref_y: np.ndarray, w: int, h: int):
    nmb = (w // 16) * (h // 16)
    mv = np.zeros((nmb, 2), np.int16)
    sad = np.zeros(nmb, np.int32)
    ev = C.c_int32()
    lib().icsp_oracle_me(_p(np.ascontiguousarray(cur_y)), _p(np.ascontiguousarray(ref_y)), w, h, _p(mv), _p(sad), C.byref(ev))
    return mv, sad, ev.value


# ---- the compiled, unmodified reference (oracle/_ref) -----------------------------------------------
def have_ref() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, b)) for b in ("ICSPCodec_O2", "ICSPDecoder_O2", "ref_taps"))


def ref_encode(frames: np.ndarray, qdc: int, qac: int, ip: int, threads: int = 0, binary: str = "ICSPCodec_O2",
               workdir: str | None = None):
    """Run the reference encoder CLI (CIF only). Returns (bin bytes or None in MT mode, recon uint8 [n][fb])."""
    n = frames.shape[0]
    tmp = workdir or tempfile.mkdtemp(prefix="icspref_")
    try:
        np.ascontiguousarray(frames, np.uint8).tofile(os.path.join(tmp, "clip_cif.yuv"))
        cmd = [os.path.join(REF_DIR, binary), "-i", "clip_cif.yuv", "-n", str(n), "--qpdc", str(qdc), "--qpac", str(qac),
               "--intraPeriod", str(ip)]
        if threads:
            cmd += ["--EnMultiThread", str(threads)]
        subprocess.run(cmd, cwd=tmp, check=True, stdout=subprocess.DEVNULL)
        recon = np.fromfile(os.path.join(tmp, "test_yuv.yuv"), np.uint8)
        binp = os.path.join(tmp, f"clip_compCIF_{qdc}_{qac}_{ip}.bin")
        data = open(binp, "rb").read() if os.path.exists(binp) else None
        return data, recon.reshape(-1, 352 * 288 * 3 // 2)
    finally:
        if workdir is None:
            shutil.rmtree(tmp, ignore_errors=True)


def ref_decode(data: bytes, nframes: int, qdc: int, qac: int, ip: int) -> np.ndarray:
    """Run the reference decoder CLI; it opens files literally named 'output\\<bin>' and 'data\\<yuv>' (DEC.h:241,323)."""
    tmp = tempfile.mkdtemp(prefix="icspref_")
    try:
        with open(os.path.join(tmp, "output\\s.bin"), "wb") as f:
            f.write(data)
        np.zeros(nframes * 352 * 288 * 3 // 2, np.uint8).tofile(os.path.join(tmp, "data\\o.yuv"))
        subprocess.run([os.path.join(REF_DIR, "ICSPDecoder_O2"), str(nframes), "s.bin", str(qdc), str(qac), str(ip), "o.yuv"],
                       cwd=tmp, check=True, stdout=subprocess.DEVNULL)
        name = "check_test_intra_yuv.yuv" if ip == 1 else "check_test_inter_yuv.yuv"
        return np.fromfile(os.path.join(tmp, name), np.uint8).reshape(-1, 352 * 288 * 3 // 2)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def ref_full_mv(frames: np.ndarray, qdc: int, qac: int, ip: int) -> np.ndarray:
    n = frames.shape[0]
    tmp = tempfile.mkdtemp(prefix="icspref_")
    try:
        np.ascontiguousarray(frames, np.uint8).tofile(os.path.join(tmp, "clip_cif.yuv"))
        subprocess.run([os.path.join(REF_DIR, "ref_taps"), "mv", "clip_cif.yuv", str(n), str(qdc), str(qac), str(ip), "mv.bin"],
                       cwd=tmp, check=True, stdout=subprocess.DEVNULL)
        return np.fromfile(os.path.join(tmp, "mv.bin"), np.int32).reshape(n, 396, 2)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def ref_dct(blocks: np.ndarray, inverse: bool = False) -> np.ndarray:
    tmp = tempfile.mkdtemp(prefix="icspref_")
    try:
        np.ascontiguousarray(blocks, np.int32).tofile(os.path.join(tmp, "in.bin"))
        subprocess.run([os.path.join(REF_DIR, "ref_taps"), "idct" if inverse else "dct", "in.bin", "out.bin"], cwd=tmp, check=True)
        return np.fromfile(os.path.join(tmp, "out.bin"), np.float64).reshape(-1, 64)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def ref_quant(dct: np.ndarray, qdc: int, qac: int, chroma: bool):
    """Quantization_block / CQuantization_block of the compiled reference on stand-alone blocks: (levels int32 [n][64], acflag int32 [n])."""
    tmp = tempfile.mkdtemp(prefix="icspref_")
    try:
        d = np.ascontiguousarray(dct, np.float64).reshape(-1, 64)
        d.tofile(os.path.join(tmp, "in.bin"))
        subprocess.run([os.path.join(REF_DIR, "ref_taps"), "quant", "in.bin", str(qdc), str(qac), "1" if chroma else "0", "out.bin"], cwd=tmp, check=True)
        o = np.fromfile(os.path.join(tmp, "out.bin"), np.int32)
        n = d.shape[0]
        return o[: n * 64].reshape(n, 64), o[n * 64:]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
