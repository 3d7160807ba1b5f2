"""CPU: libtnb.so loads, exports every symbol include/tnb.h declares, the ctypes
table binds all of them, host-only entry points answer, and the product path
refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tnb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tnb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    from tncontract_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libtnb.so does not export " + n
        assert n in _lib.SIGNATURES, "ctypes table misses " + n
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.load().tnb_version() >= 100


def test_host_only_entry_points():
    from tncontract_b200 import _lib
    lib = _lib.load()
    assert lib.tnb_error_string(0) == b"ok"
    assert b"converge" in lib.tnb_error_string(_lib.E_NOCONV)
    assert lib.tnb_norm2_workspace() > 0
    for code in (_lib.F64, _lib.C128):
        assert lib.tnb_qr_workspace(code, 3072, 1536) > 3072 * 1536 * (8 << code)
        assert lib.tnb_svd_workspace(code, 1024, 1536) > lib.tnb_qr_workspace(code, 1536, 1024)
        assert lib.tnb_qr_workspace(code, 0, 5) == 0
    # planner: a middle-axis contraction needs no permutation workspace, a scattered one does
    a = _lib.make_desc(0, _lib.C128, (2, 9, 11), (99, 11, 1))
    b = _lib.make_desc(0, _lib.C128, (4, 9), (9, 1))
    ax, bx = (ctypes.c_int32 * 1)(1), (ctypes.c_int32 * 1)(1)
    assert lib.tnb_tensordot_workspace(ctypes.byref(a), ctypes.byref(b), 1, ax, bx) == 0
    c = _lib.make_desc(0, _lib.F64, (3, 4, 5, 6), (120, 30, 6, 1))
    d = _lib.make_desc(0, _lib.F64, (4, 6, 7), (42, 7, 1))
    ax2, bx2 = (ctypes.c_int32 * 2)(1, 3), (ctypes.c_int32 * 2)(0, 1)
    assert lib.tnb_tensordot_workspace(ctypes.byref(c), ctypes.byref(d), 2, ax2, bx2) >= 3 * 4 * 5 * 6 * 8
    # argument validation happens before any CUDA call
    bad = (ctypes.c_int32 * 2)(0, 0)
    assert lib.tnb_permute(ctypes.byref(c), bad, ctypes.c_void_p(8), 1.0, 0.0, 0, None) < 0
    assert lib.tnb_gemm(7, 0, 0, 4, 4, 4, (ctypes.c_double * 2)(1, 0), None, 4, 0, None, 4, 0,
                        (ctypes.c_double * 2)(0, 0), ctypes.c_void_p(8), 4, 0, 1, None) < 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import tncontract_b200 as tn
    from tncontract_b200 import _lib
    with pytest.raises(_lib.TnbError):
        tn.Tensor(np.zeros((2, 2)), ["a", "b"])
    with pytest.raises(_lib.TnbError):
        tn.onedim.init_mps_allzero(3, 2)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tncontract_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src and "fake_tnb" not in src, os.path.join(dirpath, f)
