"""Diagonal-pattern mode (csrc/stage_pattern.cuh): operands whose entries sit on <= 64 distinct
diagonals each and whose product has <= 256.  Parity against the oracle exactly like the general
path (rowptrC / colC bit-exact, values bit-exact on integer inputs, 1e-12 / 1e-5 on reals), plus
the switch-over cases: one diagonal too many, too many output diagonals, unstructured operands."""
import numpy as np
import pytest

import oracle
from benchmark_spgemm_using_csr_b200 import BHSPARSE_CUDA, NUM_PLATFORMS, bhsparse, generators as gen, spgemm
from conftest import assert_csr_equal

pytestmark = pytest.mark.gpu
RTOL = {np.float64: 1e-12, np.float32: 1e-5}


def _oracle(A, B):
    return oracle.spgemm(A.rows, A.cols, B.cols, A.rowptr, A.col, A.val, B.rowptr, B.col, B.val)


def _check(A, B, what, pattern, exact=True):
    rp, col, val, st, prods = spgemm(A, B, return_stats=True, return_row_products=True)
    assert st["pattern_mode"] == (1 if pattern else 0), f"{what}: pattern_mode = {st['pattern_mode']}"
    assert_csr_equal((rp, col, val), _oracle(A, B), exact_values=exact, rtol=RTOL[A.val.dtype.type], what=what)
    want_p, total = oracle.row_products(A.rows, A.rowptr, A.col, B.rowptr)
    assert st["products"] == total and np.array_equal(prods.astype(np.int64), want_p)
    return st


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("name,args", [("poisson5pt", (70, 50)), ("poisson9pt", (33, 47)), ("poisson7pt", (12, 9, 14)),
                                       ("poisson27pt", (17, 11, 13))])
def test_stencils(name, args, dt):
    A = getattr(gen, name)(*args, dtype=dt)
    st = _check(A, A, f"{name}{args}", pattern=True)
    assert st["pattern_nDA"] == {"poisson5pt": 5, "poisson9pt": 9, "poisson7pt": 7, "poisson27pt": 27}[name]


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_stencil_real_values(dt):
    A = gen.poisson27pt(9, 10, 11, dtype=dt, values="real")
    _check(A, A, "27pt real", pattern=True, exact=False)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("keep", [1.0, 0.6])
def test_different_operands_irregular_rows(dt, keep):
    """A and B with different diagonal sets; keep < 1 drops entries at random, so rows of B hold
    irregular subsets of DB (the generic mask path of the symbolic kernel) and rows of A vary."""
    n = 3000
    A = gen.diagonals(n, n, [-700, -31, -2, -1, 0, 1, 5, 64, 900], keep=keep, seed=3, dtype=dt)
    B = gen.diagonals(n, n, [-1000, -65, -3, 0, 1, 2, 3, 4, 33, 512, 1024], keep=keep, seed=4, value_seed=5, dtype=dt)
    st = _check(A, B, f"diagonals keep={keep}", pattern=True)
    assert st["pattern_nDA"] == 9 and st["pattern_nDB"] == 11


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_rectangular_and_row_block(dt):
    """Rectangular operands and a row block of A (the multi-GPU partition: offsets shift by the block start)."""
    A = gen.diagonals(2500, 1800, [-400, -1, 0, 1, 7, 300], dtype=dt)
    B = gen.diagonals(1800, 2200, [-5, 0, 1, 2, 450], value_seed=9, dtype=dt)
    _check(A, B, "rectangular", pattern=True)
    S = gen.poisson27pt(14, 14, 14, dtype=dt)
    _check(S.row_slice(700, 1900), S, "row block", pattern=True)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_wide_bands(dt):
    """64 diagonals in A (the per-operand limit), B rows of up to 40 entries (two lane chunks)."""
    n = 1500
    A = gen.diagonals(n, n, list(range(-32, 32)), dtype=dt)
    B = gen.diagonals(n, n, list(range(-20, 20)), value_seed=6, dtype=dt)
    st = _check(A, B, "wide bands", pattern=True)
    assert st["pattern_nDA"] == 64 and st["pattern_nD"] == 103


def test_limits_fall_back_to_the_general_path():
    n = 2000
    A65 = gen.diagonals(n, n, list(range(-32, 33)))                  # 65 diagonals
    _check(A65, A65, "65 diagonals", pattern=False)
    wide = gen.diagonals(n, n, [17 * i for i in range(-15, 16)])     # 31 x 31 -> 61 sums: fine
    _check(wide, wide, "31 diagonals", pattern=True)
    spread = gen.diagonals(n, n, [i * i for i in range(40)])         # 40 x 40 offsets -> > 256 distinct sums
    _check(spread, spread, "too many output diagonals", pattern=False)
    R = gen.rmat(11, 8)
    _check(R, R, "unstructured", pattern=False)
    _check(R, gen.diagonals(1 << 11, 1 << 11, [-1, 0, 1]), "unstructured x banded", pattern=False)


def test_pattern_switch_and_reuse(monkeypatch):
    """BHB200_PATTERN=off gives the general path (same result); spgemm_numeric reuses the pattern plan."""
    A = gen.poisson27pt(12, 12, 12)
    want = _oracle(A, A)
    monkeypatch.setenv("BHB200_PATTERN", "off")
    rp, col, val, st = spgemm(A, A, return_stats=True)
    assert st["pattern_mode"] == 0
    assert_csr_equal((rp, col, val), want, what="pattern off")
    monkeypatch.delenv("BHB200_PATTERN")
    platforms = [False] * NUM_PLATFORMS
    platforms[BHSPARSE_CUDA] = True
    bh = bhsparse()
    assert bh.initPlatform(platforms) == 0
    rowptrC = np.zeros(A.rows + 1, dtype=np.int32)
    assert bh.initData(A.rows, A.cols, A.cols, A.nnz, A.val, A.rowptr, A.col, A.nnz, A.val, A.rowptr, A.col, rowptrC) == 0
    assert bh.spgemm() == 0 and bh.stats()["pattern_mode"] == 1
    assert bh.spgemm() == 0                                     # second call: cached plan
    v2 = gen.int_values(A.nnz, 77)
    assert bh.update_values(v2, None) == 0 and bh.spgemm_numeric() == 0
    n = bh.get_nnzC()
    col, val = np.empty(n, np.int32), np.empty(n, np.float64)
    assert bh.get_C(col, val) == 0
    want2 = oracle.spgemm(A.rows, A.cols, A.cols, A.rowptr, A.col, v2, A.rowptr, A.col, A.val)
    assert_csr_equal((rowptrC, col, val), want2, what="numeric reuse in pattern mode")
    bh.free_mem()
    bh.freePlatform()


def test_cached_plan_is_verified_on_every_call(monkeypatch):
    """A context reuses the plan of its previous product without the detection pass; every entry is
    checked against it, and operands with other diagonals (or none) fall back to the full detection /
    the general path.  Same results with BHB200_PATTERN=detect (detection on every call)."""
    seq = [gen.poisson27pt(10, 11, 12), gen.poisson27pt(10, 11, 12),        # same plan twice
           gen.poisson27pt(9, 13, 12),                                      # other strides: other offsets
           gen.poisson7pt(12, 12, 12),                                      # fewer diagonals
           gen.diagonals(1700, 1700, [-40, -1, 0, 1, 2, 40], keep=0.8),     # subset rows
           gen.rmat(10, 8),                                                 # no structure at all
           gen.poisson5pt(40, 41)]                                          # and back
    want_mode = [1, 1, 1, 1, 1, 0, 1]
    for env in (None, "detect"):
        if env:
            monkeypatch.setenv("BHB200_PATTERN", env)
        platforms = [False] * NUM_PLATFORMS
        platforms[BHSPARSE_CUDA] = True
        bh = bhsparse()
        assert bh.initPlatform(platforms) == 0
        for A, mode in zip(seq, want_mode):
            rowptrC = np.zeros(A.rows + 1, dtype=np.int32)
            assert bh.initData(A.rows, A.cols, A.cols, A.nnz, A.val, A.rowptr, A.col, A.nnz, A.val, A.rowptr, A.col, rowptrC) == 0
            assert bh.spgemm() == 0, bh.last_error()
            assert bh.stats()["pattern_mode"] == mode
            n = bh.get_nnzC()
            col, val = np.empty(n, np.int32), np.empty(n, np.float64)
            assert bh.get_C(col, val) == 0
            assert_csr_equal((rowptrC, col, val), _oracle(A, A), what=f"sequence step ({env})")
        # A and B with different diagonal sets after a square product (the cached lists must not be mixed up)
        A = gen.diagonals(900, 900, [-3, 0, 5])
        B = gen.diagonals(900, 900, [-7, -1, 0, 2, 11], value_seed=4)
        rowptrC = np.zeros(A.rows + 1, dtype=np.int32)
        assert bh.initData(A.rows, A.cols, B.cols, A.nnz, A.val, A.rowptr, A.col, B.nnz, B.val, B.rowptr, B.col, rowptrC) == 0
        assert bh.spgemm() == 0 and bh.stats()["pattern_mode"] == 1
        n = bh.get_nnzC()
        col, val = np.empty(n, np.int32), np.empty(n, np.float64)
        assert bh.get_C(col, val) == 0
        assert_csr_equal((rowptrC, col, val), _oracle(A, B), what="A != B after A == B")
        assert bh.spgemm() == 0                         # same operands again: speculative hit
        assert bh.get_C(col, val) == 0
        assert_csr_equal((rowptrC, col, val), _oracle(A, B), what="A != B again")
        bh.free_mem()
        bh.freePlatform()
