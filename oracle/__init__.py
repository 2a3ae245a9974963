"""ctypes front-end of the CPU oracle (oracle/spgemm_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(benchmark_spgemm_using_csr_b200) never imports this module.
"""
import ctypes
import os

import numpy as np

from . import build as _build

_LIB = None

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def lib():
    global _LIB
    if _LIB is None:
        path = _build.OUT
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(_build.SRC):
            path = _build.build()
        L = ctypes.CDLL(path)
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_set_num_threads.argtypes = [ctypes.c_int]
        L.oracle_row_products.restype = ctypes.c_int64
        L.oracle_row_products.argtypes = [ctypes.c_int, _i32p, _i32p, _i32p, _i64p]
        L.oracle_reference_bin.restype = ctypes.c_int
        L.oracle_reference_bin.argtypes = [ctypes.c_int64]
        L.oracle_spgemm_symbolic.restype = ctypes.c_int64
        L.oracle_spgemm_symbolic.argtypes = [ctypes.c_int, ctypes.c_int, _i32p, _i32p, _i32p, _i32p, _i64p]
        for name, vp in (("f64", _f64p), ("f32", _f32p)):
            f = getattr(L, "oracle_spgemm_numeric_" + name)
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_int, ctypes.c_int, _i32p, _i32p, vp, _i32p, _i32p, vp, _i64p, _i32p, vp]
            g = getattr(L, "oracle_csr_sort_indices_" + name)
            g.restype = ctypes.c_int
            g.argtypes = [ctypes.c_int, _i32p, _i32p, vp]
        _LIB = L
    return _LIB


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(t: int) -> None:
    lib().oracle_set_num_threads(int(t))


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def row_products(m, rowptrA, colA, rowptrB):
    """Per-row intermediate-product counts (int64[m]) and their total."""
    out = np.zeros(max(m, 1), dtype=np.int64)
    total = lib().oracle_row_products(m, _c(rowptrA, np.int32), _c(colA, np.int32), _c(rowptrB, np.int32), out)
    return out[:m], int(total)


def reference_bin(count: int) -> int:
    return int(lib().oracle_reference_bin(int(count)))


def spgemm(m, k, n, rowptrA, colA, valA, rowptrB, colB, valB):
    """C = A*B.  Returns (rowptrC int64[m+1], colC int32[nnzC], valC[nnzC]) with
    the dtype of valA (float32 or float64); columns ascending per row, explicit
    zeros kept."""
    del k
    valA = np.ascontiguousarray(valA)
    if valA.dtype not in (np.float32, np.float64):
        raise TypeError("values must be float32 or float64")
    dt = valA.dtype
    rowptrA, colA = _c(rowptrA, np.int32), _c(colA, np.int32)
    rowptrB, colB, valB = _c(rowptrB, np.int32), _c(colB, np.int32), _c(valB, dt)
    rowptrC = np.zeros(m + 1, dtype=np.int64)
    nnzC = lib().oracle_spgemm_symbolic(m, n, rowptrA, colA, rowptrB, colB, rowptrC)
    if nnzC < 0:
        raise MemoryError("oracle symbolic phase failed")
    colC = np.empty(max(nnzC, 1), dtype=np.int32)
    valC = np.empty(max(nnzC, 1), dtype=dt)
    fn = lib().oracle_spgemm_numeric_f64 if dt == np.float64 else lib().oracle_spgemm_numeric_f32
    if fn(m, n, rowptrA, colA, valA, rowptrB, colB, valB, rowptrC, colC, valC) != 0:
        raise MemoryError("oracle numeric phase failed")
    return rowptrC, colC[:nnzC], valC[:nnzC]


def csr_sort_indices(rows, rowptr, col, val):
    """In-place per-row sort by column (ref_spgemm.h:37-62)."""
    fn = lib().oracle_csr_sort_indices_f64 if val.dtype == np.float64 else lib().oracle_csr_sort_indices_f32
    if fn(rows, _c(rowptr, np.int32), col, val) != 0:
        raise MemoryError("oracle sort failed")
