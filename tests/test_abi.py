"""not gpu: the C-ABI shared library builds for sm_100a, loads, and exports every symbol include/b200gan.h declares."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "b200gan.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    lib = ctypes.CDLL(ge.LIB)
    names = _declared()
    assert len(names) >= 38
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    lib.b200_version.restype = ctypes.c_int
    assert lib.b200_version() == 100
    lib.b200_last_error.restype = ctypes.c_char_p
    assert lib.b200_last_error() is not None


def test_library_is_blackwell_native():
    """SASS carries tcgen05 MMA (UTCHMMA), TMEM loads (LDTM) and TMA loads (UTMALDG) — B200_PROFILING.md evidence table"""
    import __graft_entry__ as ge
    ge.build()
    sass = subprocess.run(["cuobjdump", "-sass", ge.LIB], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", ge.LIB], capture_output=True, text=True).stdout


def test_product_fails_loudly_without_gpu():
    import pytest
    import torch
    from b200gan import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    k = _lib.Kernels()
    with pytest.raises(_lib.B200Error):
        k.relu_fwd(torch.zeros(8))
