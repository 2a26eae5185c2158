"""The C-ABI library builds, loads, and exports every symbol include/stedm_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from tests.util import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "stedm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(stedm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_header_symbols():
    from stedm_b200 import build, _lib
    path = build.build()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    lib.stedm_abi_version.restype = ctypes.c_int
    assert lib.stedm_abi_version() == 3


def test_conv_desc_layout_matches_header():
    from stedm_b200._lib import ConvDesc
    assert ctypes.sizeof(ConvDesc) == 9 * 8 + 8 + 22 * 4 + 2 * 8 + 6 * 4 + 8  # 9 ptrs, int64, 22 int32, 2 ptrs, 6 int32, 1 ptr


def test_sass_is_blackwell_native():
    """tcgen05.mma / TMA / tcgen05.ld must be in the shipped SASS (UTCHMMA / UTMALDG / LDTM)."""
    import shutil
    import subprocess
    from stedm_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", build.build()], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
