"""The C-ABI library must load without a GPU and export every symbol include/neunet_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "numpy-nn-model_b200", "lib", "libneunet_b200.so")
HDR = os.path.join(ROOT, "include", "neunet_b200.h")


def _declared():
    text = open(HDR).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nnb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def dll():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(LIB)


def test_header_declares_the_path():
    names = _declared()
    for must in ["nnb_linear_forward", "nnb_linear_backward", "nnb_matmul_forward", "nnb_matmul_backward",
                 "nnb_conv2d_forward", "nnb_conv2d_backward", "nnb_swish_forward", "nnb_softmax_forward",
                 "nnb_rmsnorm_forward", "nnb_adamw_step"]:
        assert must in names


def test_every_declared_symbol_is_exported(dll):
    missing = [n for n in _declared() if not hasattr(dll, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_python_binding_matches_header(dll):
    from neunet import b200
    assert sorted(b200.EXPORTED_SYMBOLS) == _declared()
    b200.lib()  # resolves every symbol with its signature; raises on mismatch


def test_no_compute_entry_points(dll):
    dll.nnb_version.restype = ctypes.c_int
    assert dll.nnb_version() >= 100
    dll.nnb_last_error.restype = ctypes.c_char_p
    assert isinstance(dll.nnb_last_error(), bytes)
    dll.nnb_weight_staged_bytes.restype = ctypes.c_size_t
    dll.nnb_weight_staged_bytes.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
    assert dll.nnb_weight_staged_bytes(128, 784, 0) == 128 * 784 * 2
    assert dll.nnb_weight_staged_bytes(10, 10, 1) == 2 * 512  # 10 rows x ld 16 x 2 B = 320 -> 512 per plane, two planes
    # argument validation happens before any CUDA call
    dll.nnb_linear_forward.restype = ctypes.c_int
    assert dll.nnb_linear_forward(None, None, None, None, None, ctypes.c_int64(1), ctypes.c_int64(1), ctypes.c_int64(1),
                                  0, ctypes.c_float(1), 0, None, None, None, ctypes.c_size_t(0), None) == 1
    assert b"null" in dll.nnb_last_error()


def test_product_never_imports_oracle():
    """A product path routed through the oracle would void every parity claim."""
    pkg = os.path.join(ROOT, "numpy-nn-model_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"(import\s+oracle|from\s+oracle|oracle[./\\]|oracle\s*import)", src), \
                    f"{f} references the oracle package"
