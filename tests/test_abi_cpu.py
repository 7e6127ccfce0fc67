"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/icspcuda.h declares,
and fails loudly (no fallback) when no CUDA device is present."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from icspcodec_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from icspcodec_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "icspcuda.h")).read()
    declared = set(re.findall(r"\b(icsp_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_version_and_bad_params(lib):
    import ctypes as C
    assert b"sm_100a" in lib.icsp_version()
    h = C.c_void_p()
    assert lib.icsp_create(C.byref(h), 0, 350, 288, 4) == -1      # width not a multiple of 16
    assert b"multiples of 16" in lib.icsp_last_error(None)
    assert lib.icsp_create(C.byref(h), 0, 352, 288, 0) == -1


def test_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from icspcodec_b200 import IcspCuda, IcspError
    with pytest.raises(IcspError, match="no CPU fallback"):
        IcspCuda(352, 288, 4)
