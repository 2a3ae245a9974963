"""Python mirror of the reference's `bhsparse` class (SpGEMM_cuda/bhsparse.h:17-34)
over the C-ABI: same method names, argument order (val, rowptr, colidx) and
`int` error convention (0 == BHSPARSE_SUCCESS, common.h:26), so that the parity
tests read like the reference driver (main.cu:104-135).

    bh = bhsparse()
    bh.initPlatform(platforms)          # platforms[BHSPARSE_CUDA] = True
    bh.initData(m, k, n, nnzA, valA, rowptrA, colA, nnzB, valB, rowptrB, colB, rowptrC)
    bh.warmup(); bh.spgemm()
    nnzC = bh.get_nnzC(); bh.get_C(colC, valC)
    bh.free_mem(); bh.freePlatform()

Host arrays are numpy (or anything exposing a writable buffer of the right
dtype); the value dtype (float32/float64) is taken from valA -- the reference
needs a recompile for that (common.h:31, README.md:84-86).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import capi

BHSPARSE_SUCCESS = 0
NUM_PLATFORMS = 9       # common.h:33
BHSPARSE_CUDA = 1       # common.h:36


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


class bhsparse:
    def __init__(self, device: int = 0, verbose: bool = False):
        self._lib = None
        self._ctx = ctypes.c_void_p(None)
        self._device = device
        self._verbose = verbose
        self._rowptrC = None
        self._dtype = None
        self._m = 0

    # -- platform (bhsparse.h:96-148) ---------------------------------------
    def initPlatform(self, spgemm_platform) -> int:
        if len(spgemm_platform) < NUM_PLATFORMS or not spgemm_platform[BHSPARSE_CUDA]:
            return capi.ERR_INVALID          # only the CUDA platform exists here
        self._lib = capi.load()
        ctx = ctypes.c_void_p(None)
        err = self._lib.bhb200_create(ctypes.byref(ctx), self._device)
        if err == BHSPARSE_SUCCESS:
            self._ctx = ctx
            if self._verbose:               # bhsparse_cuda.h:108-110
                print(f"Device [{self._device}] {self._lib.bhb200_device_name(ctx).decode()}. "
                      f"{self._lib.bhb200_sm_count(ctx)} SMs.")
        return err

    def freePlatform(self) -> int:
        if not self._ctx:
            return BHSPARSE_SUCCESS
        err = self._lib.bhb200_destroy(self._ctx)
        self._ctx = ctypes.c_void_p(None)
        return err

    # -- data (bhsparse.h:180-258) -------------------------------------------
    def initData(self, m, k, n, nnzA, csrValA, csrRowPtrA, csrColIndA,
                 nnzB, csrValB, csrRowPtrB, csrColIndB, csrRowPtrC) -> int:
        if not self._ctx:
            return capi.ERR_INVALID
        valA = np.ascontiguousarray(csrValA)
        if valA.dtype == np.float64:
            fn, dt = self._lib.bhb200_init_data_f64, np.float64
        elif valA.dtype == np.float32:
            fn, dt = self._lib.bhb200_init_data_f32, np.float32
        else:
            return capi.ERR_INVALID
        valB = np.ascontiguousarray(csrValB, dtype=dt)
        rpA = np.ascontiguousarray(csrRowPtrA, dtype=np.int32)
        cA = np.ascontiguousarray(csrColIndA, dtype=np.int32)
        rpB = np.ascontiguousarray(csrRowPtrB, dtype=np.int32)
        cB = np.ascontiguousarray(csrColIndB, dtype=np.int32)
        if rpA.size != m + 1 or rpB.size != k + 1 or cA.size < nnzA or cB.size < nnzB:
            return capi.ERR_INVALID
        if csrRowPtrC is not None and (csrRowPtrC.dtype != np.int32 or csrRowPtrC.size < m + 1):
            return capi.ERR_INVALID
        self._rowptrC = csrRowPtrC           # caller-owned, written by get_C (main.cu:100)
        self._dtype = dt
        self._m = m
        return fn(self._ctx, m, k, n, nnzA, _ptr(valA), _ptr(rpA), _ptr(cA), nnzB, _ptr(valB), _ptr(rpB), _ptr(cB))

    def aliased_operands(self) -> bool:
        """True if initData got the same host arrays for A and B and uploaded them once."""
        return bool(self._ctx) and bool(self._lib.bhb200_operands_aliased(self._ctx))

    def warmup(self) -> int:
        return self._lib.bhb200_warmup(self._ctx) if self._ctx else capi.ERR_INVALID

    # -- the hot path (bhsparse.h:260-339) -------------------------------------
    def spgemm(self) -> int:
        if not self._ctx:
            return capi.ERR_INVALID
        err = self._lib.bhb200_spgemm(self._ctx)
        # create_C leaves rowptrC in the caller's array at the end of spgemm() (bhsparse_cuda.h:2787-2808)
        if err == BHSPARSE_SUCCESS and self._rowptrC is not None and 0 <= self.get_nnzC() <= 0x7fffffff:
            fn = self._lib.bhb200_get_C_f64 if self._dtype == np.float64 else self._lib.bhb200_get_C_f32
            err = fn(self._ctx, ctypes.c_void_p(self._rowptrC.ctypes.data), None, None)
        if err == BHSPARSE_SUCCESS and self._verbose:
            st = self.stats()
            t = st["ms_total"]
            gf = 2.0 * st["products"] / (t * 1.0e6) if t > 0 else 0.0
            print(f"STAGE 1 time: {st['ms_count']:.4f} ms.\nSTAGE 2 time: {st['ms_symbolic']:.4f} ms.\n"
                  f"STAGE 3 time: {st['ms_scan']:.4f} ms.\nSTAGE 4 time: {st['ms_numeric']:.4f} ms.\n"
                  f"[ CUDA ] SpGEMM time: {t:.4f} ms. Gflops = {gf:.4f}")    # bhsparse.h:286-289
        return err

    def get_nnzC(self) -> int:
        return int(self._lib.bhb200_get_nnzC(self._ctx)) if self._ctx else -1

    def get_C(self, csrColIndC, csrValC) -> int:
        if not self._ctx:
            return capi.ERR_INVALID
        if csrValC is not None and csrValC.dtype != self._dtype:
            return capi.ERR_INVALID
        if csrColIndC is not None and csrColIndC.dtype != np.int32:
            return capi.ERR_INVALID
        fn = self._lib.bhb200_get_C_f64 if self._dtype == np.float64 else self._lib.bhb200_get_C_f32
        rp = ctypes.c_void_p(self._rowptrC.ctypes.data) if self._rowptrC is not None else None
        cp = ctypes.c_void_p(csrColIndC.ctypes.data) if csrColIndC is not None else None
        vp = ctypes.c_void_p(csrValC.ctypes.data) if csrValC is not None else None
        return fn(self._ctx, rp, cp, vp)

    def free_mem(self) -> int:
        return self._lib.bhb200_free_mem(self._ctx) if self._ctx else BHSPARSE_SUCCESS

    # -- additions (no reference counterpart) --------------------------------------
    def update_values(self, csrValA=None, csrValB=None) -> int:
        """New values for A and/or B, same patterns (include/bhsparse_b200.h: bhb200_update_values_*)."""
        if not self._ctx:
            return capi.ERR_INVALID
        for v in (csrValA, csrValB):
            if v is not None and (v.dtype != self._dtype or not v.flags.c_contiguous):
                return capi.ERR_INVALID
        fn = self._lib.bhb200_update_values_f64 if self._dtype == np.float64 else self._lib.bhb200_update_values_f32
        ap = ctypes.c_void_p(csrValA.ctypes.data) if csrValA is not None else None
        bp = ctypes.c_void_p(csrValB.ctypes.data) if csrValB is not None else None
        return fn(self._ctx, ap, bp)

    def spgemm_numeric(self) -> int:
        """Values of C again after update_values; structure of C reused (bhb200_spgemm_numeric)."""
        return self._lib.bhb200_spgemm_numeric(self._ctx) if self._ctx else capi.ERR_INVALID

    def get_rowptrC_i64(self) -> np.ndarray:
        out = np.empty(self._m + 1, dtype=np.int64)
        capi.check(self._lib, self._ctx, self._lib.bhb200_get_rowptrC_i64(self._ctx, ctypes.c_void_p(out.ctypes.data)))
        return out

    def get_C_range(self, first: int, count: int):
        """Entries [first, first+count) of (colC, valC): piecewise access for results beyond INT32_MAX."""
        col = np.empty(max(count, 0), dtype=np.int32)
        val = np.empty(max(count, 0), dtype=self._dtype)
        capi.check(self._lib, self._ctx, self._lib.bhb200_get_C_range(
            self._ctx, int(first), int(count), ctypes.c_void_p(col.ctypes.data), ctypes.c_void_p(val.ctypes.data)))
        return col, val

    def get_row_products(self) -> np.ndarray:
        out = np.empty(max(self._m, 1), dtype=np.int32)
        capi.check(self._lib, self._ctx, self._lib.bhb200_get_row_products(self._ctx, ctypes.c_void_p(out.ctypes.data)))
        return out[:self._m]

    def stats(self) -> dict:
        st = capi.Stats()
        capi.check(self._lib, self._ctx, self._lib.bhb200_get_stats(self._ctx, ctypes.byref(st)))
        return st.as_dict()

    def last_error(self) -> str:
        return self._lib.bhb200_last_error(self._ctx).decode() if self._ctx else ""

    def __del__(self):
        try:
            if self._ctx:
                self.freePlatform()
        except Exception:
            pass


def spgemm(A, B, device: int = 0, return_stats: bool = False, return_row_products: bool = False):
    """Convenience driver following main.cu:104-135: C = A*B for two
    generators.CSR operands; returns (rowptrC int32, colC int32, valC)."""
    platforms = [False] * NUM_PLATFORMS
    platforms[BHSPARSE_CUDA] = True
    bh = bhsparse(device)
    rowptrC = np.zeros(A.rows + 1, dtype=np.int32)

    def ok(err, what):
        if err != BHSPARSE_SUCCESS:
            msg = bh.last_error()
            bh.freePlatform()
            raise capi.BhsparseError(err, f"{what}: {msg}")

    ok(bh.initPlatform(platforms), "initPlatform")
    ok(bh.initData(A.rows, A.cols, B.cols, A.nnz, A.val, A.rowptr, A.col,
                   B.nnz, B.val, B.rowptr, B.col, rowptrC), "initData")
    ok(bh.spgemm(), "spgemm")
    nnzC = bh.get_nnzC()
    colC = np.empty(max(nnzC, 0), dtype=np.int32)
    valC = np.empty(max(nnzC, 0), dtype=A.val.dtype)
    ok(bh.get_C(colC, valC), "get_C")
    st = bh.stats() if return_stats else None
    prods = bh.get_row_products() if return_row_products else None
    ok(bh.free_mem(), "free_mem")
    ok(bh.freePlatform(), "freePlatform")
    out = (rowptrC, colC, valC)
    if return_stats:
        out += (st,)
    if return_row_products:
        out += (prods,)
    return out
