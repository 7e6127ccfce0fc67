"""icspcodec_b200 — B200-native (sm_100a) implementation of ICSPCodec's encode/decode core.

Product = icspcodec_b200/libicspcuda.so (C ABI, include/icspcuda.h) + the C++ host tools in
icspcodec_b200/host (icspenc, icspdec).  The Python modules are bindings for tests and bench.py.
"""
from .api import EncResult, IcspCuda, IcspError, PinnedArray, finish_stream, stream_header  # noqa: F401
