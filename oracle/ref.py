"""ctypes front-end of oracle/_ref -- the reference itself (bhSPARSE CUDA), compiled by
oracle/build_ref.py.  TEST INFRASTRUCTURE ONLY: tests/, smoke() and bench.py's reference legs.

    rowptrC, colC, valC, ms = ref.spgemm(m, k, n, rowptrA, colA, valA, rowptrB, colB, valB)

runs main.cu:104-135's call protocol on device 0 (hard-wired in the reference,
bhsparse_cuda.h:100-101) and returns the reference's own C plus the wall time of its
spgemm() call (bhsparse.h:268-289: stages 1-4, allocations included).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import build_ref as _build_ref

_LIBS: dict = {}


def available() -> bool:
    return len(_build_ref.available()) == 2


def _lib(kind: str):
    if kind not in _LIBS:
        paths = _build_ref.available()
        if kind not in paths:
            raise FileNotFoundError("oracle/_ref is not built (python -m oracle.build_ref, needs /root/reference)")
        L = ctypes.CDLL(paths[kind])
        L.bhref_value_size.restype = ctypes.c_int
        L.bhref_log.restype = ctypes.c_char_p
        L.bhref_cuda_errors.restype = ctypes.c_int
        L.bhref_spgemm.restype = ctypes.c_int
        L.bhref_spgemm.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p] * 3 + [ctypes.c_int] + [ctypes.c_void_p] * 4 + \
                                  [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)]
        L.bhref_get_C.restype = ctypes.c_int
        L.bhref_get_C.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        assert L.bhref_value_size() == (8 if kind == "f64" else 4)
        _LIBS[kind] = L
    return _LIBS[kind]


class ReferenceError_(RuntimeError):
    pass


def spgemm(m, k, n, rowptrA, colA, valA, rowptrB, colB, valB, warmups: int = 0, fetch: bool = True):
    valA = np.ascontiguousarray(valA)
    dt = valA.dtype
    kind = "f64" if dt == np.float64 else "f32"
    L = _lib(kind)
    rowptrA = np.ascontiguousarray(rowptrA, dtype=np.int32)
    colA = np.ascontiguousarray(colA, dtype=np.int32)
    rowptrB = np.ascontiguousarray(rowptrB, dtype=np.int32)
    colB = np.ascontiguousarray(colB, dtype=np.int32)
    valB = np.ascontiguousarray(valB, dtype=dt)
    rowptrC = np.zeros(m + 1, dtype=np.int32)
    nnzC = ctypes.c_int(0)
    ms = ctypes.c_double(0.0)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    err = L.bhref_spgemm(m, k, n, colA.size, p(valA), p(rowptrA), p(colA), colB.size, p(valB), p(rowptrB), p(colB),
                         p(rowptrC), warmups, ctypes.byref(nnzC), ctypes.byref(ms))
    if err != 0:
        L.bhref_get_C(None, None)
        raise ReferenceError_(f"reference spgemm failed, code {err}; log:\n{L.bhref_log().decode(errors='replace')[-2000:]}")
    nz = int(nnzC.value)
    colC = np.empty(max(nz, 1), dtype=np.int32)
    valC = np.empty(max(nz, 1), dtype=dt)
    err = L.bhref_get_C(p(colC), p(valC)) if fetch else L.bhref_get_C(None, None)
    if err != 0:
        raise ReferenceError_(f"reference get_C failed, code {err}")
    return rowptrC, colC[:nz], valC[:nz], float(ms.value)


def log(kind: str = "f64") -> str:
    return _lib(kind).bhref_log().decode(errors="replace")
