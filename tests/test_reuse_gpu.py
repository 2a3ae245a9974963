"""Repeated products with the same patterns: bhb200_update_values_* +
bhb200_spgemm_numeric must give the oracle's C for the NEW values, bit for bit on the
structure, on every kernel family (direct-mode rows, wide bins, CTA tables, global bitmap,
range kernels, ESC), in both precisions."""
import numpy as np
import pytest

import oracle
from benchmark_spgemm_using_csr_b200 import BHSPARSE_CUDA, BHSPARSE_SUCCESS, NUM_PLATFORMS, bhsparse, capi, generators as gen
from conftest import assert_csr_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["general", "default"])
def _path(request, monkeypatch):
    """Every test of this module runs twice: with the diagonal-pattern mode switched off (the
    general hash / ESC / range / bitmap kernels these tests were written for) and with the
    library's default (structured operands then take csrc/stage_pattern.cuh)."""
    if request.param == "general":
        monkeypatch.setenv("BHB200_PATTERN", "off")


def _new_values(nnz, seed, dt):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(1, 10, size=nnz).astype(dt)


def _run(A, B, what, rounds=2):
    dt = A.val.dtype.type
    platforms = [False] * NUM_PLATFORMS
    platforms[BHSPARSE_CUDA] = True
    bh = bhsparse()
    assert bh.initPlatform(platforms) == BHSPARSE_SUCCESS
    rowptrC = np.zeros(A.rows + 1, dtype=np.int32)
    assert bh.initData(A.rows, A.cols, B.cols, A.nnz, A.val, A.rowptr, A.col, B.nnz, B.val, B.rowptr, B.col,
                       rowptrC) == BHSPARSE_SUCCESS
    assert bh.spgemm_numeric() == capi.ERR_INVALID          # nothing to reuse yet
    assert bh.spgemm() == BHSPARSE_SUCCESS
    nnzC = bh.get_nnzC()
    for r in range(rounds):
        va = _new_values(A.nnz, 100 + r, dt)
        vb = _new_values(B.nnz, 200 + r, dt) if r % 2 == 0 else None      # second round: only A changes
        assert bh.update_values(va, vb) == BHSPARSE_SUCCESS
        assert bh.spgemm_numeric() == BHSPARSE_SUCCESS
        assert bh.get_nnzC() == nnzC
        colC = np.empty(nnzC, dtype=np.int32)
        valC = np.empty(nnzC, dtype=dt)
        assert bh.get_C(colC, valC) == BHSPARSE_SUCCESS
        if vb is not None:
            cur_b = vb
        want = oracle.spgemm(A.rows, A.cols, B.cols, A.rowptr, A.col, va, B.rowptr, B.col, cur_b)
        assert_csr_equal((rowptrC, colC, valC), want, exact_values=True, rtol=0, what=f"{what} round {r}")
        st = bh.stats()
        assert st["num_bin_rows"][17] == 0 and st["ms_symbolic"] < 0.05     # no copy bin, no symbolic pass
    # a full product afterwards still works and agrees
    assert bh.spgemm() == BHSPARSE_SUCCESS
    colC2 = np.empty(nnzC, dtype=np.int32)
    valC2 = np.empty(nnzC, dtype=dt)
    assert bh.get_C(colC2, valC2) == BHSPARSE_SUCCESS
    assert np.array_equal(colC, colC2) and np.array_equal(valC, valC2)
    bh.free_mem()
    bh.freePlatform()


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_reuse_stencil_direct_rows(dt, monkeypatch):
    monkeypatch.setenv("BHB200_RANGE", "off")      # hash path: the rows run in direct mode on the full product
    A = gen.poisson27pt(24, 24, 24, dtype=dt)
    _run(A, A, f"27pt reuse {dt.__name__}")


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_reuse_stencil_range_rows(dt):
    A = gen.poisson27pt(20, 20, 20, dtype=dt)       # narrow spans: bitmap-rank kernels + word lists
    _run(A, A, f"27pt range reuse {dt.__name__}")
    A = gen.poisson5pt(64, 64, dtype=dt)            # ESC rows
    _run(A, A, f"5pt reuse {dt.__name__}")


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_reuse_rmat(dt):
    A = gen.rmat(13, 16, a=0.57, b=0.19, c=0.19, d=0.05, seed=3, dtype=dt)     # CTA tables + global bitmap rows
    _run(A, A, f"rmat13 reuse {dt.__name__}", rounds=1)
    A = gen.rmat(16, 8, seed=5, dtype=dt)                                      # wide direct bins on the full product
    _run(A, A, f"rmat16 reuse {dt.__name__}", rounds=1)


def test_update_values_rejects_borrowed_and_wrong_type():
    A = gen.poisson5pt(16, 16)
    platforms = [False] * NUM_PLATFORMS
    platforms[BHSPARSE_CUDA] = True
    bh = bhsparse()
    bh.initPlatform(platforms)
    rowptrC = np.zeros(A.rows + 1, dtype=np.int32)
    bh.initData(A.rows, A.cols, A.cols, A.nnz, A.val, A.rowptr, A.col, A.nnz, A.val, A.rowptr, A.col, rowptrC)
    assert bh.update_values(A.val.astype(np.float32), None) == capi.ERR_INVALID
    assert bh.update_values(None, None) == BHSPARSE_SUCCESS
    bh.freePlatform()


def test_reuse_empty_product():
    """A without entries: the numeric-only call must succeed and leave an empty C."""
    A = gen.CSR(5, 7, np.zeros(6, dtype=np.int32), np.zeros(0, dtype=np.int32), np.zeros(0))
    B = gen.poisson5pt(7, 1)
    platforms = [False] * NUM_PLATFORMS
    platforms[BHSPARSE_CUDA] = True
    bh = bhsparse()
    bh.initPlatform(platforms)
    rowptrC = np.full(6, -1, dtype=np.int32)
    assert bh.initData(5, 7, 7, 0, A.val, A.rowptr, A.col, B.nnz, B.val, B.rowptr, B.col, rowptrC) == BHSPARSE_SUCCESS
    assert bh.spgemm() == BHSPARSE_SUCCESS and bh.get_nnzC() == 0
    assert bh.update_values(None, B.val * 2) == BHSPARSE_SUCCESS
    assert bh.spgemm_numeric() == BHSPARSE_SUCCESS and bh.get_nnzC() == 0
    assert bh.get_C(np.empty(0, dtype=np.int32), np.empty(0)) == BHSPARSE_SUCCESS
    assert np.array_equal(rowptrC, np.zeros(6, dtype=np.int32))
    bh.freePlatform()
