"""ctypes loader for libicspcuda.so.  Fails loudly: there is no CPU or PyTorch fallback for this path."""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libicspcuda.so")

ICSP_OK = 0
MAX_KERNELS = 32


class EncOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("levels", "acflag", "mpm", "ipm", "mvd", "mv", "minsad", "recon", "dct")]


class DecIn(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("levels", "mpm", "ipm", "mvd")]


class DecBitsIn(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("bits", "stream_offset", "stream_bytes", "row_bit_offset")]


class BitsOut(C.Structure):
    _fields_ = [("bits", C.c_void_p), ("cap_bytes", C.c_size_t), ("stream_bits", C.c_void_p), ("stream_offset", C.c_void_p),
                ("recon", C.c_void_p)]


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 40), ("launches", C.c_uint64), ("total_ms", C.c_double)]


# every symbol include/icspcuda.h declares (tests check that all of them are exported)
EXPORTS = [
    "icsp_create", "icsp_destroy", "icsp_last_error", "icsp_version", "icsp_sync",
    "icsp_encode_gops", "icsp_enc_upload", "icsp_enc_run", "icsp_enc_download",
    "icsp_decode_gops", "icsp_dec_upload", "icsp_dec_run", "icsp_dec_download",
    "icsp_me_sad", "icsp_dct8x8", "icsp_idct8x8",
    "icsp_set_profiling", "icsp_reset_stats", "icsp_get_stats", "icsp_launch_count",
    "icsp_event_record", "icsp_event_elapsed_ms",
    "icsp_host_alloc", "icsp_host_alloc_upload", "icsp_host_free",
    "icsp_configure", "icsp_encode_streams", "icsp_entropy_run", "icsp_bits_download", "icsp_finish_body", "icsp_bits_bound", "icsp_enc_sse", "icsp_bits_row_index", "icsp_decode_streams",
    "icsp_quant", "icsp_intra_frame", "icsp_inter_frame",
]

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m icspcodec_b200.build` "
                           "(libicspcuda has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i = C.c_void_p, C.c_int
    lib.icsp_create.argtypes = [C.POINTER(vp), i, i, i, i]
    lib.icsp_destroy.argtypes = [vp]
    lib.icsp_destroy.restype = None
    lib.icsp_last_error.argtypes = [vp]
    lib.icsp_last_error.restype = C.c_char_p
    lib.icsp_version.restype = C.c_char_p
    lib.icsp_sync.argtypes = [vp]
    lib.icsp_encode_gops.argtypes = [vp, vp, i, i, i, i, C.POINTER(EncOut)]
    lib.icsp_enc_upload.argtypes = [vp, vp, i]
    lib.icsp_enc_run.argtypes = [vp, i, i, i, i]
    lib.icsp_enc_download.argtypes = [vp, i, C.POINTER(EncOut)]
    lib.icsp_enc_sse.argtypes = [vp, i, vp]
    lib.icsp_bits_row_index.argtypes = [vp, i, vp]
    lib.icsp_decode_streams.argtypes = [vp, C.POINTER(DecBitsIn), i, i, i, i, i, vp]
    lib.icsp_decode_gops.argtypes = [vp, C.POINTER(DecIn), i, i, i, i, vp]
    lib.icsp_dec_upload.argtypes = [vp, C.POINTER(DecIn), i]
    lib.icsp_dec_run.argtypes = [vp, i, i, i, i]
    lib.icsp_dec_download.argtypes = [vp, i, vp]
    lib.icsp_me_sad.argtypes = [vp, vp, vp, i, vp, vp]
    lib.icsp_dct8x8.argtypes = [vp, vp, i, vp]
    lib.icsp_idct8x8.argtypes = [vp, vp, i, i, vp]
    lib.icsp_quant.argtypes = [vp, vp, i, i, i, i, vp, vp]
    lib.icsp_intra_frame.argtypes = [vp, vp, i, i, i, C.POINTER(EncOut)]
    lib.icsp_inter_frame.argtypes = [vp, vp, vp, i, i, i, C.POINTER(EncOut)]
    lib.icsp_set_profiling.argtypes = [vp, i]
    lib.icsp_reset_stats.argtypes = [vp]
    lib.icsp_get_stats.argtypes = [vp, C.POINTER(KernelStat), i]
    lib.icsp_launch_count.argtypes = [vp]
    lib.icsp_launch_count.restype = C.c_uint64
    lib.icsp_configure.argtypes = [vp, i, i]
    lib.icsp_event_record.argtypes = [vp, i]
    lib.icsp_event_elapsed_ms.argtypes = [vp, i, i, C.POINTER(C.c_float)]
    lib.icsp_encode_streams.argtypes = [vp, vp, i, i, i, i, i, C.POINTER(BitsOut)]
    lib.icsp_entropy_run.argtypes = [vp, i, i, i]
    lib.icsp_bits_download.argtypes = [vp, i, C.POINTER(BitsOut)]
    lib.icsp_finish_body.argtypes = [vp, C.c_uint64]
    lib.icsp_finish_body.restype = C.c_size_t
    lib.icsp_bits_bound.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.icsp_bits_bound.restype = C.c_size_t
    lib.icsp_host_alloc.argtypes = [C.c_size_t]
    lib.icsp_host_alloc.restype = vp
    lib.icsp_host_alloc_upload.argtypes = [C.c_size_t]
    lib.icsp_host_alloc_upload.restype = vp
    lib.icsp_host_free.argtypes = [vp]
    lib.icsp_host_free.restype = None
    _lib = lib
    return lib
