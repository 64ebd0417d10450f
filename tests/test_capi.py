"""CPU: the C-ABI shared library loads and exports every symbol include/smslu.h declares; host-only
entry points behave; numeric entry points fail loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import HAVE_GPU, ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "smslu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smslu_[a-z_]+)\s*\(", src)))


def test_header_symbols_exported(smslu):
    from sharedmemsparselu_jl_b200 import _capi
    lib = C.CDLL(_capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_capi.EXPORTS) == names
    assert lib.smslu_version() >= 100


def test_struct_sizes_match_header(smslu):
    from sharedmemsparselu_jl_b200 import _capi
    assert C.sizeof(_capi.Options) == 4 * (1 + 3 + 6 + 8)
    assert C.sizeof(_capi.Stats) == 8 * (14 + 9 + 5 + 16 + 16 + 8)


def test_create_rejects_bad_input(smslu):
    from sharedmemsparselu_jl_b200 import _capi
    L = _capi.lib()
    h = C.c_void_p()
    colptr = np.array([0, 1, 2], np.int64); rowval = np.array([0, 5], np.int64)
    assert L.smslu_create(C.byref(h), 2, colptr.ctypes.data, rowval.ctypes.data, 0, None) == 0
    assert L.smslu_analyze(h, None, None) == _capi.E_PATTERN          # row index out of range
    assert b"row index" in L.smslu_last_error(h)
    L.smslu_destroy(h)
    assert L.smslu_create(C.byref(h), 0, colptr.ctypes.data, rowval.ctypes.data, 0, None) == _capi.E_ARG
    assert L.smslu_create(C.byref(h), 2, colptr.ctypes.data, rowval.ctypes.data, 7, None) == _capi.E_ARG


def test_one_based_input_equals_zero_based(smslu, W):
    from sharedmemsparselu_jl_b200 import _capi
    L = _capi.lib()
    A = W.laplacian_2d(9)
    n = A.shape[0]
    outs = []
    for base in (0, 1):
        h = C.c_void_p()
        cp = (A.indptr + base).astype(np.int64); rv = (A.indices + base).astype(np.int64)
        assert L.smslu_create(C.byref(h), n, cp.ctypes.data, rv.ctypes.data, base, None) == 0
        assert L.smslu_analyze(h, None, None) == 0
        p = np.zeros(n, np.int64); q = np.zeros(n, np.int64)
        lp = np.zeros(n + 1, np.int64)
        assert L.smslu_get_factors(h, lp.ctypes.data, None, None, None, None, None, p.ctypes.data, q.ctypes.data, None, base) == 0
        outs.append((p - base, lp - base))
        L.smslu_destroy(h)
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_python_mirror_argument_errors(smslu, W):
    import scipy.sparse as sp
    with pytest.raises(TypeError):
        smslu.ParallelSparseLU(W.laplacian_2d(4).tocsr())
    with pytest.raises(smslu.DimensionMismatch):
        smslu.ParallelSparseLU(sp.csc_matrix(np.ones((2, 3))))


@pytest.mark.skipif(HAVE_GPU, reason="only meaningful without a GPU")
def test_numeric_path_fails_loudly_without_gpu(smslu, W):
    """No CPU fallback: without a device the constructor raises SMSLU_E_CUDA."""
    from sharedmemsparselu_jl_b200 import _capi
    with pytest.raises(smslu.SmsluError) as ei:
        smslu.ParallelSparseLU(W.laplacian_2d(6))
    assert ei.value.code == _capi.E_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/."""
    pkg = os.path.join(ROOT, "sharedmemsparselu.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".jl")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and False, os.path.join(dirpath, f)


def test_duplicate_entries_are_rejected():
    """A non-canonical CSC with a repeated (row, column) would silently lose a value in the A -> front scatter."""
    import ctypes as C
    import numpy as np
    from sharedmemsparselu_jl_b200 import _capi
    L = _capi.lib()
    colptr = np.array([0, 3, 5], np.int64); rowval = np.array([0, 1, 1, 0, 1], np.int64)     # (1,0) twice
    h = C.c_void_p()
    assert L.smslu_create(C.byref(h), 2, colptr.ctypes.data_as(C.c_void_p), rowval.ctypes.data_as(C.c_void_p), 0, None) == 0
    assert L.smslu_analyze(h, None, None) == _capi.E_PATTERN
    assert b"duplicate" in L.smslu_last_error(h)
    L.smslu_destroy(h)
