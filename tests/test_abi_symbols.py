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


def test_argument_errors_of_loss_and_training_entry_points():
    """error convention of the entry points added for the loss / epilogue / training rows: 0 for empty inputs,
    PZ_ERR_ARG (-1) for null pointers and bad sizes, PZ_ERR_UNSUPPORTED (-2) for shapes outside the supported range;
    all decided before the first CUDA call (safe without a GPU)."""
    lib = _lib.load()
    one = ctypes.c_void_p(16)            # a non-null, 16-byte aligned dummy: only reached by checks that pass first
    assert lib.pz_chamfer(None, None, 0, 8, 8, None, None, None, None, None) == 0
    assert lib.pz_chamfer(None, None, 2, 8, 8, None, None, None, None, None) == -1
    assert lib.pz_chamfer(one, one, 2, 0, 8, one, one, None, None, None) == -1       # min over an empty set
    assert b"empty" in lib.pz_last_error()
    assert lib.pz_comp(one, one, 0, one, None) == -1
    assert lib.pz_boundary_topk(one, 1, 2048, 128, one, None, None) == -2
    assert lib.pz_boundary_topk(one, 1, 1024, 2000, one, None, None) == -1
    assert lib.pz_topk(one, 1, 0, 1, 1, one, None, None) == -2
    assert lib.pz_topk(None, 0, 10, 1, 1, None, None, None) == 0
    assert lib.pz_pair_score(None, None, None, None, None, None, None, None, None, None, 0, None, None, None, None,
                             None, None) == 0
    assert lib.pz_pair_score(None, None, None, None, None, None, None, None, None, None, 3, None, None, None, None,
                             None, None) == -1
    assert lib.pz_se3_transform(None, None, 2, 0, None, None) == 0
    # pz_sgemm: split-K excludes epilogues and batching
    assert lib.pz_sgemm(0, 0, 8, 8, 8, 1.0, one, 8, one, 8, 0.0, one, 8, 1, 0, 0, 0, 4, one, 0, None, 0, None, 0, None) == -2
    assert lib.pz_sgemm(0, 0, 8, 8, 0, 1.0, one, 8, one, 8, 0.0, one, 8, 1, 0, 0, 0, 1, None, 0, None, 0, None, 0, None) == -1
    assert lib.pz_sgemm(0, 0, 0, 8, 8, 1.0, None, 8, None, 8, 0.0, None, 8, 1, 0, 0, 0, 1, None, 0, None, 0, None, 0, None) == 0
    # pz_gemm_tf32: shape and alignment contract
    assert lib.pz_gemm_tf32(0, 0, 100, 128, 32, one, 32, one, 32, one, 128, 1, None, 0, None, 0, 0, None) == -2
    assert lib.pz_gemm_tf32(0, 0, 128, 96, 32, one, 32, one, 32, one, 96, 1, None, 0, None, 0, 0, None) == -2
    assert b"M % 128" in lib.pz_last_error()
    assert lib.pz_gemm_tf32(0, 0, 128, 128, 32, one, 30, one, 32, one, 128, 1, None, 0, None, 0, 0, None) == -1
    assert lib.pz_gemm_tf32(0, 0, 128, 128, 32, one, 32, one, 32, one, 128, 4, one, 0, None, 0, 0, None) == -2
    assert lib.pz_bn_point_train_forward(one, 2, 4, 4, one, one, one, None, 0.1, 1e-5, 1, one, one, one, None) == -1
    assert lib.pz_maxpool_forward(None, 0, 4, 4, None, None, None) == 0
    assert lib.pz_maxpool_forward(one, 2, 0, 4, one, one, None) == -1
    assert lib.pz_cross_entropy(one, one, 0, 8, 0, 1.0, one, None, None) == -1
    assert lib.pz_adam_step(one, one, one, one, 8, 1e-3, 0.9, 0.999, 1e-8, 0, 1.0, None) == -1      # step >= 1
    assert lib.pz_pose_grad(one, one, None, 8, None, 0.0, 2, 0.0, one, None) == -1                  # pts without dpts
