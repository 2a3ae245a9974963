"""BASELINE.json's five configs at FULL size, entry for entry against the CPU oracle
(rowptrC and colC bit-exact; values bit-exact -- the inputs are the driver's integers 1..9,
main.cu:82,93, so every sum is exact in f32 and f64).

  C1 Poisson5pt 1024^2 f64        C2 Poisson27pt 128^3 f64       C4 uniform rect 4Mx1M * 1Mx4M f32
  C3 R-MAT scale 22 ef16, f64 AND f32: nnz(C) = 2.53e9 > INT32_MAX, so the result is read through
     the int64 row pointers + bhb200_get_C_range and compared row block by row block (the oracle
     runs on the same row block of A), which also bounds host memory;
  C5's partition: R-MAT scale 21 split into 2 / 4 / 8 row blocks on the prefix sum of the per-row
     products (dist.partition_rows_by_products), each block multiplied separately and the blocks
     re-assembled with their nnz(C) offsets -- the multi-GPU scheme run on one device (the real
     N-rank run is tests/test_dist_gpu.py and bench.py's parity gate).
Size-independent properties (checksum of checksums, sortedness) are kept as a second check."""
import numpy as np
import pytest

import oracle
from benchmark_spgemm_using_csr_b200 import (BHSPARSE_CUDA, NUM_PLATFORMS, bhsparse, capi, generators as gen, spgemm)
from benchmark_spgemm_using_csr_b200.dist import partition_rows_by_products, row_cost, row_products_host
from conftest import assert_csr_equal

pytestmark = pytest.mark.gpu


def _oracle(A, B):
    return oracle.spgemm(A.rows, A.cols, B.cols, A.rowptr, A.col, A.val, B.rowptr, B.col, B.val)


def _properties(A, B, rp, col, val, st, P=None, nnzC=None):
    assert rp[0] == 0 and rp[-1] == col.size == val.size == st["nnzC"]
    if P is not None:
        assert st["products"] == P
    if nnzC is not None:
        assert st["nnzC"] == nnzC
    # checksum of checksums: sum_ij C_ij = sum_k colsum(A)_k * rowsum(B)_k
    colsumA = np.bincount(A.col, weights=A.val.astype(np.float64), minlength=A.cols)
    rowsumB = np.add.reduceat(B.val.astype(np.float64), B.rowptr[:-1].astype(np.int64)) * (np.diff(B.rowptr) > 0)
    want = float(np.dot(colsumA, rowsumB))
    got = float(val.astype(np.float64).sum())
    assert abs(got - want) <= 1e-9 * abs(want)


def _full_compare(A, B, what, P=None, nnzC=None):
    rp, col, val, st = spgemm(A, B, return_stats=True)
    _properties(A, B, rp, col, val, st, P, nnzC)
    assert_csr_equal((rp, col, val), _oracle(A, B), exact_values=True, what=what)
    prods, total = oracle.row_products(A.rows, A.rowptr, A.col, B.rowptr)
    assert st["products"] == total
    return st


def test_config1_poisson5pt_1024():
    A = gen.poisson5pt(1024, 1024)
    _full_compare(A, A, "C1", P=26177544, nnzC=13611012)


def test_config2_poisson27pt_128():
    A = gen.poisson27pt(128, 128, 128)
    _full_compare(A, A, "C2", P=1489355288, nnzC=254840104)


def test_config4_uniform_rect_f32():
    A = gen.uniform_rect(4194304, 1048576, per_row=8, seed=1, dtype=np.float32)
    B = gen.uniform_rect(1048576, 4194304, per_row=8, seed=2, value_seed=3, dtype=np.float32)
    _full_compare(A, B, "C4", P=268435456)


def test_config2_and_config4_real_values_within_tolerance():
    """one extra run per config with uniform reals -- the tolerance path at full size
    (BASELINE.json: <= 1e-12 relative in double, <= 1e-5 in float; structure bit-exact)."""
    A = gen.poisson27pt(128, 128, 128)
    A = gen.CSR(A.rows, A.cols, A.rowptr, A.col, gen.real_values(A.col.size, 11, np.float64))
    rp, col, val = spgemm(A, A)
    assert_csr_equal((rp, col, val), _oracle(A, A), exact_values=False, rtol=1e-12, what="C2 real values")
    A4 = gen.uniform_rect(4194304, 1048576, per_row=8, seed=1, dtype=np.float32)
    B4 = gen.uniform_rect(1048576, 4194304, per_row=8, seed=2, value_seed=3, dtype=np.float32)
    A4 = gen.CSR(A4.rows, A4.cols, A4.rowptr, A4.col, gen.real_values(A4.col.size, 12, np.float32))
    B4 = gen.CSR(B4.rows, B4.cols, B4.rowptr, B4.col, gen.real_values(B4.col.size, 13, np.float32))
    rp, col, val = spgemm(A4, B4)
    assert_csr_equal((rp, col, val), _oracle(A4, B4), exact_values=False, rtol=1e-5, what="C4 real values")


@pytest.fixture(scope="module")
def rmat22():
    return gen.rmat(22, 16)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_config3_rmat22_int64_blockwise(rmat22, dt):
    A = rmat22.astype(dt)
    platforms = [False] * NUM_PLATFORMS
    platforms[BHSPARSE_CUDA] = True
    bh = bhsparse()
    assert bh.initPlatform(platforms) == 0
    rowptrC32 = np.zeros(A.rows + 1, dtype=np.int32)
    assert bh.initData(A.rows, A.cols, A.cols, A.nnz, A.val, A.rowptr, A.col, A.nnz, A.val, A.rowptr, A.col, rowptrC32) == 0
    assert bh.spgemm() == 0, bh.last_error()
    st = bh.stats()
    assert st["nnzC"] > 0x7fffffff and bh.get_nnzC() == st["nnzC"]
    # the int32 API refuses, the int64 one serves
    assert bh.get_C(np.empty(1, np.int32), np.empty(1, dt)) == capi.ERR_OVERFLOW
    rp64 = bh.get_rowptrC_i64()
    assert rp64[0] == 0 and rp64[-1] == st["nnzC"]
    prods, total = oracle.row_products(A.rows, A.rowptr, A.col, A.rowptr)
    assert st["products"] == total
    assert np.array_equal(bh.get_row_products().astype(np.int64), prods)
    nblocks = 16
    bounds = partition_rows_by_products(prods, nblocks)
    for b in range(nblocks):
        r0, r1 = int(bounds[b]), int(bounds[b + 1])
        if r1 == r0:
            continue
        blk = A.row_slice(r0, r1)
        wrp, wcol, wval = _oracle(blk, A)
        assert np.array_equal(rp64[r0:r1 + 1] - rp64[r0], wrp), f"C3 block {b}: rowptrC differs"
        col, val = bh.get_C_range(int(rp64[r0]), int(rp64[r1] - rp64[r0]))
        assert np.array_equal(col, wcol), f"C3 block {b}: colC differs"
        assert np.array_equal(val, wval), f"C3 block {b}: valC differs"
    assert bh.free_mem() == 0 and bh.freePlatform() == 0


@pytest.fixture(scope="module")
def rmat21():
    A = gen.rmat(21, 16)
    return A, _oracle(A, A)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_config5_partition_rmat21(rmat21, world):
    """R-MAT scale 21 (config 5's generator, three scales down) through the N-way row-block scheme."""
    A, (wrp, wcol, wval) = rmat21
    prods = row_products_host(A, A.rowptr)
    cost = row_cost(prods)
    bounds = partition_rows_by_products(cost, world)
    share = [int(cost[bounds[r]:bounds[r + 1]].sum()) for r in range(world)]
    assert max(share) <= 1.05 * (sum(share) / world) + cost.max()        # blocks balanced on the cost of their products
    off = 0
    for r in range(world):
        r0, r1 = int(bounds[r]), int(bounds[r + 1])
        rp, col, val = spgemm(A.row_slice(r0, r1), A)
        assert np.array_equal(rp.astype(np.int64) + off, wrp[r0:r1 + 1]), f"rank {r}: rowptrC + offset differs"
        assert np.array_equal(col, wcol[off:off + col.size]) and np.array_equal(val, wval[off:off + col.size])
        off += int(rp[-1])
    assert off == int(wrp[-1])
