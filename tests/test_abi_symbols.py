"""CPU: libpuzzlenet_sm100.so loads and exports exactly what include/puzzlenet_b200.h declares, and the
ctypes binding covers every declared function (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

from puzzlenet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "puzzlenet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(pz_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_is_valid_c():
    import subprocess
    subprocess.check_call(["gcc", "-fsyntax-only", "-x", "c", "-Wall", "-Werror",
                           os.path.join(ROOT, "include", "puzzlenet_b200.h")])


def test_library_exports_every_declared_symbol():
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_binding_matches_header():
    assert sorted(_lib.SIGNATURES) == declared_functions()
    lib = _lib.load()
    assert lib.pz_abi_version() == _lib.ABI_VERSION


def test_argument_errors_without_gpu():
    lib = _lib.load()
    # null pointers are rejected before any CUDA call, so this is safe on a CPU-only box
    assert lib.pz_fps(None, 1, 16, None, 4, None, None, None) == -1
    assert b"null" in lib.pz_last_error()
    assert lib.pz_knn(None, None, 1, 1, 1, 1, None, None, None) == -1
    assert lib.pz_knn(None, None, 0, 0, 1, 1, None, None, None) == 0          # empty input is a no-op
    assert lib.pz_predict5_workspace_bytes(64) > 0
    assert lib.pz_encoder_workspace_bytes(2, 64) > lib.pz_encoder_workspace_bytes(1, 64)
    with pytest.raises((RuntimeError, ValueError)):
        _lib.call("pz_sqdist", None, None, 1, 1, 1, None, None)
