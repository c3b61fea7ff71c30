"""The C-ABI library loads (no GPU needed) and exports every symbol that
include/stylish_b200.h declares; the ctypes table covers the same set."""
import ctypes
import os
import re

from stylish_tts_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "stylish_b200.h")).read()
    return sorted(set(re.findall(r"STY_API\s+[\w\s\*]+?\b(sty_\w+)\s*\(", src)))


def test_library_exports_header_symbols():
    from stylish_tts_b200.csrc import build

    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.EXPORTED) == names


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.sty_version() >= 100
    assert isinstance(lib.sty_last_error(), bytes)


def test_bad_arguments_are_rejected_without_gpu():
    lib = _lib.load()
    a = _lib.ConvArgs()  # all-null
    rc = lib.sty_conv1d_fwd(ctypes.byref(a), None)
    assert rc == -1
    assert b"null" in lib.sty_last_error()
